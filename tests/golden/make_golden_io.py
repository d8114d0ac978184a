#!/usr/bin/env python
"""Golden FILES of the output side, written by the REFERENCE'S OWN Python
(imported from /root/reference, nothing copied): `utils/io.py` (PFM, camera text),
`utils/eval_file_logger.py` (the per-view file set of test.py:76) and the Gipuma `.dmb` writer of
`tools/depthfusion.py`.  Run once in the build container; `io_golden.npz` holds the inputs and the
bytes of every file the reference wrote.

Adjustments, applied from outside, for APIs numpy 2 removed: `np.int` / `np.float`
(eval_file_logger.py:55,100) are restored as aliases of the builtins, and arrays are handed to the
reference's `write_pfm` as an ndarray subclass that still answers `.tostring()` (io.py:142) with
`.tobytes()`.
"""
import io as _io
import os
import runpy
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

np.int = int      # noqa: E305  (see module docstring)
np.float = float

from pointmvsnet.utils import io as ref_io  # noqa: E402
import pointmvsnet.utils.eval_file_logger as ref_logger_mod  # noqa: E402


class _Arr(np.ndarray):
    def tostring(self):
        return self.tobytes()


def ref_write_pfm(path, image, scale=1):
    return ref_io.write_pfm(path, image.view(_Arr), scale)


ref_logger_mod.write_pfm = ref_write_pfm
ref_logger = ref_logger_mod.eval_file_logger


def file_bytes(path):
    with open(path, "rb") as f:
        return np.frombuffer(f.read(), dtype=np.uint8)


def main():
    g = torch.Generator().manual_seed(7)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        # ---- PFM / cam / dmb primitives ---------------------------------------------------------
        img = torch.randn(6, 5, generator=g).numpy().astype(np.float32)
        ref_write_pfm(os.path.join(tmp, "a.pfm"), img, scale=1)
        out["pfm_image"] = img
        out["pfm_bytes"] = file_bytes(os.path.join(tmp, "a.pfm"))
        back, scale = ref_io.load_pfm(os.path.join(tmp, "a.pfm"))
        assert scale == 1.0 and np.array_equal(back, img)

        cam = np.zeros((2, 4, 4), dtype=np.float32)
        cam[0] = torch.randn(4, 4, generator=g).numpy()
        cam[0, 3] = (0, 0, 0, 1)
        cam[1, :3, :3] = [[361.54, 0.0, 82.9], [0.0, 360.4, 66.4], [0.0, 0.0, 1.0]]
        cam[1, 3] = (425.0, 2.5, 48.0, 542.5)
        ref_io.write_cam_dtu(os.path.join(tmp, "cam.txt"), cam)
        out["cam"] = cam
        out["cam_bytes"] = file_bytes(os.path.join(tmp, "cam.txt"))
        text = open(os.path.join(tmp, "cam.txt")).read()
        out["cam_loaded_31"] = ref_io.load_cam_dtu(_io.StringIO(text), num_depth=96, interval_scale=0.5)
        words = text.split()
        out["cam_loaded_29"] = ref_io.load_cam_dtu(_io.StringIO(" ".join(words[:29])), num_depth=96, interval_scale=0.5)
        out["cam_loaded_30"] = ref_io.load_cam_dtu(_io.StringIO(" ".join(words[:30])), num_depth=96, interval_scale=0.5)
        out["cam_loaded_27"] = ref_io.load_cam_dtu(_io.StringIO(" ".join(words[:27])), num_depth=96, interval_scale=0.5)

        tools = runpy.run_path(os.path.join(REF, "tools", "depthfusion.py"), run_name="ref_depthfusion")
        tools["write_gipuma_dmb"](os.path.join(tmp, "d.dmb"), img)
        out["dmb_bytes"] = file_bytes(os.path.join(tmp, "d.dmb"))
        assert np.array_equal(tools["read_gipuma_dmb"](os.path.join(tmp, "d.dmb")), img)
        img3 = torch.randn(4, 3, 3, generator=g).numpy().astype(np.float32)
        tools["write_gipuma_dmb"](os.path.join(tmp, "d3.dmb"), img3)
        out["dmb3_image"] = img3
        out["dmb3_bytes"] = file_bytes(os.path.join(tmp, "d3.dmb"))

        # ---- the per-view file set ---------------------------------------------------------------
        H, W = 16, 20  # "image"; flow1 at 1/4, flow2 at 1/2
        ref_img = (torch.rand(1, H, W, 3, generator=g) * 255).floor()
        cams = torch.zeros(1, 2, 2, 4, 4)
        cams[0, 0] = torch.from_numpy(cam)
        cams[0, 1] = torch.from_numpy(cam) * 1.01
        preds = {
            "coarse_depth_map": 425 + 100 * torch.rand(1, 1, H // 4, W // 4, generator=g),
            "coarse_prob_map": torch.rand(1, 1, H // 4, W // 4, generator=g),
            "flow1": 425 + 100 * torch.rand(1, 1, H // 4, W // 4, generator=g),
            "flow1_prob": torch.softmax(3 * torch.randn(1, 5, H // 4, W // 4, generator=g), dim=1),
            "flow2": 425 + 100 * torch.rand(1, 1, H // 2, W // 2, generator=g),
            "flow2_prob": torch.softmax(3 * torch.randn(1, 5, H // 2, W // 2, generator=g), dim=1),
        }
        # exercise the negative-floor wrap and the ceil clamp of eval_file_logger.py:55-58
        preds["flow1_prob"][0, :, 0, 0] = torch.tensor([1.0000001, 0.0, 0.0, 0.0, 0.0])
        preds["flow1_prob"][0, :, 0, 1] = torch.tensor([0.0, 0.0, 0.0, 0.0, 1.0])
        data_batch = {"cam_params_list": cams, "ref_img": ref_img}
        ref_path = os.path.join(tmp, "Eval", "Rectified", "scan9", "rect_001_3_r5000.png")
        ref_logger(data_batch, preds, ref_path, "out_b200")
        scene = os.path.join(tmp, "Eval", "out_b200", "scan9")
        names = sorted(os.listdir(scene))
        out["logger_names"] = np.array(names)
        for n in names:
            out["logger_file_" + n] = file_bytes(os.path.join(scene, n))
        # ---- hand-over to the fusion stage (tools/depthfusion.py:64-170) on that scene -----------------
        import cv2
        tools["probability_filter"].__globals__["write_pfm"] = ref_write_pfm  # .tostring(), see docstring
        tools["probability_filter"](scene, 0.2, 0.1, "flow2", 1, cv2.INTER_LANCZOS4)
        point_folder = os.path.join(tmp, "points", "scan9")
        os.makedirs(point_folder)
        tools["mvsnet_to_gipuma"](scene, point_folder, "flow2", 1)
        out["fusion_file_00000000_flow2_prob_filtered.pfm"] = file_bytes(
            os.path.join(scene, "00000000_flow2_prob_filtered.pfm"))
        fusion = []
        for root, _dirs, files in sorted(os.walk(point_folder)):
            for fn in sorted(files):
                rel = os.path.relpath(os.path.join(root, fn), point_folder)
                fusion.append(rel)
                out["fusion_file_" + rel] = file_bytes(os.path.join(root, fn))
        out["fusion_names"] = np.array(fusion)
        out["logger_ref_img"] = ref_img.numpy()
        out["logger_cams"] = cams.numpy()
        for k, v in preds.items():
            out["logger_pred_" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "io_golden.npz"), **out)
    print("wrote io_golden.npz:", ", ".join(names))


if __name__ == "__main__":
    main()
