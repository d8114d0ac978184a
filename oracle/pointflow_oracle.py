"""CPU oracle for the PointFlow hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
(``pointmvsnet_b200``) never imports it and has no CPU fallback.

It is a plain fp32 PyTorch-on-CPU restatement of the reference algorithm
(callmeray/PointMVSNet @ cacb2d7).  Every function cites the reference
file:line it follows.  Three documented adjustments make it equal to "the
reference's own CUDA ops" (SURVEY.md section 8c):

  1. bilinear sampling uses ``align_corners=True`` -- the reference was written
     for PyTorch 1.0.1 where that was the only behaviour, and its grid
     normalisation (utils/feature_fetcher.py:51-53) only inverts under it;
  2. EdgeConv gathers the conv2 output (the CUDA branch, networks.py:26-28),
     not the conv1 output the divergent CPU branch uses (networks.py:29-33);
  3. BatchNorm always uses batch statistics, because the reference runs
     inference under ``model.train()`` (test.py:58).

Parity pin: ``tests/golden/make_golden.py`` runs the *reference's own Python*
(imported from /root/reference, with exactly those three adjustments applied by
monkey-patching, no source copied) and stores input/output vectors under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this oracle against
them, including the reference's own known-answer test
(utils/feature_fetcher.py:63-97) and gather test (functions/gather_knn.py:27-56).

kNN order: ``torch.topk`` tie order is implementation defined, so the oracle
uses the canonical order (distance ascending, then candidate id ascending);
see ``knn_compare`` for the comparison rule used against the reference/CUDA.
"""
import math

import torch
import torch.nn.functional as F

HYPOTHESES = (-2, -1, 0, 1, 2)  # model.py:172 interval_list
BN_EPS = 1e-5  # torch.nn.BatchNorm default, networks.py:16, nn/conv.py:24


# --------------------------------------------------------------------------- #
# a3: pixel grid  (functions/functions.py:128-138)
# --------------------------------------------------------------------------- #
def get_pixel_grids(height, width):
    """[3, H*W] rows (x+0.5, y+0.5, 1), row-major over (y, x)."""
    xs = torch.linspace(0.5, width - 0.5, width)
    ys = torch.linspace(0.5, height - 0.5, height)
    gx = xs.view(1, width).expand(height, width).reshape(-1)
    gy = ys.view(height, 1).expand(height, width).reshape(-1)
    return torch.stack([gx, gy, torch.ones(height * width)], dim=0)


# --------------------------------------------------------------------------- #
# a6: FeatureFetcher.forward  (utils/feature_fetcher.py:13-60)
# --------------------------------------------------------------------------- #
def feature_fetch(feature_maps, pts, cam_intrinsics, cam_extrinsics):
    """feature_maps [B,V,C,H,W], pts [B,3,N], K [B,V,3,3], E [B,V,3,4] or None
    -> [B,V,C,N].  Projection feature_fetcher.py:29-49, grid :51-53,
    sampling :55 (bilinear, zeros padding, align_corners=True semantics)."""
    B, V, C, H, W = feature_maps.shape
    N = pts.shape[2]
    fm = feature_maps.reshape(B * V, C, H, W)
    K = cam_intrinsics.reshape(B * V, 3, 3)
    p = pts.unsqueeze(1).expand(B, V, 3, N).reshape(B * V, 3, N)
    if cam_extrinsics is None:
        cam = p.float().transpose(1, 2)
    else:
        E = cam_extrinsics.reshape(B * V, 3, 4)
        R = E[:, :, :3]
        t = E[:, :, 3:4].expand(B * V, 3, N)
        cam = (torch.bmm(R, p) + t).float().transpose(1, 2)
    x, y, z = cam[..., 0], cam[..., 1], cam[..., 2]
    nuv = torch.stack([x / z, y / z, torch.ones_like(x)], dim=-1)
    uv = torch.bmm(nuv, K.transpose(1, 2))[:, :, :2]
    grid = (uv - 0.5).view(B * V, N, 1, 2).clone()
    grid[..., 0] = (grid[..., 0] / float(W - 1)) * 2 - 1.0
    grid[..., 1] = (grid[..., 1] / float(H - 1)) * 2 - 1.0
    out = F.grid_sample(fm, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    return out.squeeze(3).view(B, V, C, N)


# --------------------------------------------------------------------------- #
# a10: get_knn_3d  (utils/torch_utils.py:16-61)
# --------------------------------------------------------------------------- #
def knn3d_dist2(xyz, kernel_size=5):
    """Squared distances to the kernel_size^3 window candidates, [B, k^3, N].

    Follows torch_utils.py:29-47: the conv3d with one-hot (+centre, -tap)
    filters and zero padding (:44) yields per axis ``centre - neighbour`` with a
    single rounding (all other taps multiply exact zeros); an out-of-grid
    neighbour is the zero vector; candidate order is d*k*k + h*k + w (:32-38);
    the squared sum runs over x, y, z in that order (:46-47)."""
    B, _, D, H, W = xyz.shape
    hk = kernel_size // 2
    padded = F.pad(xyz, (hk, hk, hk, hk, hk, hk))
    cols = []
    for dd in range(kernel_size):
        for dh in range(kernel_size):
            for dw in range(kernel_size):
                nb = padded[:, :, dd:dd + D, dh:dh + H, dw:dw + W]
                diff = xyz - nb
                sq = diff * diff
                cols.append(((sq[:, 0] + sq[:, 1]) + sq[:, 2]).reshape(B, -1))
    return torch.stack(cols, dim=1)


def knn3d(xyz, kernel_size=5, knn=16, return_dist=False):
    """xyz [B,3,D,H,W] -> int64 [B, D*H*W, knn] in canonical order
    (dist2 ascending, candidate id ascending).  Index arithmetic and the global
    clamp follow torch_utils.py:49-59."""
    B, _, D, H, W = xyz.shape
    hk = kernel_size // 2
    k2 = kernel_size * kernel_size
    dist2 = knn3d_dist2(xyz, kernel_size)
    order = torch.sort(dist2, dim=1, stable=True).indices[:, :knn]  # [B,knn,N]
    cand = order.permute(0, 2, 1)  # [B,N,knn]
    d_off = cand // k2 - hk
    h_off = (cand % k2) // kernel_size - hk
    w_off = cand % kernel_size - hk
    n = torch.arange(D * H * W).view(1, -1, 1)
    idx = n + d_off * (H * W) + h_off * W + w_off
    idx = torch.clamp(idx, 0, D * H * W - 1)
    if return_dist:
        return idx, cand, dist2
    return idx


def knn_compare(idx_test, cand_test, dist2, knn=16):
    """Comparison rule for kNN parity.

    ``dist2`` [B,125,N] are the oracle distances, ``cand_test`` [B,N,k] the
    candidate ids (0..124) picked by the implementation under test and
    ``idx_test`` its linear indices.  Returns a dict with
      tie_free_frac  : fraction of points whose 17 smallest distances are distinct
      exact_tie_free : all tie-free points have identical candidate lists
      dist_multiset  : on every point the sorted picked distances are identical
    """
    B, C, N = dist2.shape
    srt = torch.sort(dist2, dim=1, stable=True)
    top = srt.values[:, :knn + 1]  # [B,k+1,N]
    tie_free = (top[:, 1:] != top[:, :-1]).all(dim=1)  # [B,N]
    ref_cand = srt.indices[:, :knn].permute(0, 2, 1)
    same = (ref_cand == cand_test).all(dim=2)
    picked = torch.gather(dist2, 1, cand_test.permute(0, 2, 1))  # [B,k,N]
    multiset_ok = (torch.sort(picked, dim=1).values == srt.values[:, :knn]).all(dim=1)
    return {
        "tie_free_frac": tie_free.float().mean().item(),
        "exact_tie_free": bool(same[tie_free].all()),
        "exact_frac": same.float().mean().item(),
        "dist_multiset": bool(multiset_ok.all()),
    }


def idx_to_candidates(idx, D, H, W, kernel_size=5):
    """Best-effort inverse of the index arithmetic (used only to analyse
    reference output that carries no candidate ids): returns candidate ids for
    picks that were not altered by the global clamp, else -1."""
    hk = kernel_size // 2
    B, N, K = idx.shape
    n = torch.arange(N).view(1, N, 1)
    delta = idx - n
    out = torch.full_like(idx, -1)
    for dd in range(-hk, hk + 1):
        for dh in range(-hk, hk + 1):
            for dw in range(-hk, hk + 1):
                off = dd * H * W + dh * W + dw
                cid = (dd + hk) * kernel_size * kernel_size + (dh + hk) * kernel_size + (dw + hk)
                out = torch.where(delta == off, torch.full_like(out, cid), out)
    return out


# --------------------------------------------------------------------------- #
# a13: gather_knn forward  (functions/csrc/gather_knn_kernel.cu:25-47)
# --------------------------------------------------------------------------- #
def gather_knn(feature, index):
    """feature [B,C,N], index [B,N,K] int64 -> [B,C,N,K]; the reference's CUDA
    forward is an expand + torch.gather (gather_knn_kernel.cu:41-44)."""
    B, C, N = feature.shape
    K = index.shape[2]
    src = feature.unsqueeze(2).expand(B, C, N, N)
    ind = index.unsqueeze(1).expand(B, C, N, K)
    return torch.gather(src, 3, ind)


def gather_knn_backward(grad_output, index):
    """grad_output [B,C,N,K] -> grad_input [B,C,N], scatter-add
    (gather_knn_kernel.cu:50-89)."""
    B, C, N, K = grad_output.shape
    grad = torch.zeros(B, C, N, dtype=grad_output.dtype)
    ind = index.unsqueeze(1).expand(B, C, N, K).reshape(B, C, N * K)
    grad.scatter_add_(2, ind, grad_output.reshape(B, C, N * K))
    return grad


# --------------------------------------------------------------------------- #
# BatchNorm with batch statistics (test.py:58 keeps the model in train mode)
# --------------------------------------------------------------------------- #
def batch_norm_train(x, gamma, beta, eps=BN_EPS):
    """x [B,C,...]: biased batch variance over every dim but 1."""
    dims = [0] + list(range(2, x.dim()))
    mean = x.mean(dim=dims, keepdim=True)
    var = x.var(dim=dims, unbiased=False, keepdim=True)
    shape = [1, -1] + [1] * (x.dim() - 2)
    return (x - mean) / torch.sqrt(var + eps) * gamma.view(shape) + beta.view(shape)


def conv1x1(x, weight):
    """nn.Conv1d(kernel 1, bias=False): weight [Cout,Cin,1], x [B,Cin,N]."""
    return torch.matmul(weight[:, :, 0].unsqueeze(0), x)


# --------------------------------------------------------------------------- #
# a11/a12: EdgeConvNoC / EdgeConv  (networks.py:9-81, CUDA branches)
# --------------------------------------------------------------------------- #
def edge_conv(feature, knn_inds, w1, w2, gamma, beta, concat_central, eps=BN_EPS):
    """feature [B,Cin,N], knn_inds [B,N,K].
    concat_central=True : EdgeConv   (networks.py:18-45)  -> [B,2*Cout,N]
    concat_central=False: EdgeConvNoC (networks.py:56-81) -> [B,Cout,N]"""
    K = knn_inds.shape[2]
    local = conv1x1(feature, w1)  # networks.py:22 / :60
    edge = conv1x1(feature, w2)  # networks.py:23 / :61
    neighbour = gather_knn(edge, knn_inds)  # networks.py:28 / :66
    central = local.unsqueeze(-1).expand(-1, -1, -1, K)
    if concat_central:
        e = torch.cat([central, neighbour - central], dim=1)  # networks.py:39
    else:
        e = neighbour - central  # networks.py:76
    e = batch_norm_train(e, gamma, beta, eps)
    e = F.relu(e)
    return e.mean(dim=3)


# --------------------------------------------------------------------------- #
# a14: flow_mlp  (model.py:40-43; nn/mlp.py:45-81; nn/conv.py:7-41)
# --------------------------------------------------------------------------- #
def flow_mlp(x, params):
    """x [B,224,N] -> [B,1,N].  params: dict with mlp{i}_w, mlp{i}_gamma,
    mlp{i}_beta for i in 0..2 and mlp3_w (Conv1d 16->1, no bias, no BN)."""
    for i in range(3):
        x = conv1x1(x, params["mlp%d_w" % i])
        x = batch_norm_train(x, params["mlp%d_gamma" % i], params["mlp%d_beta" % i])
        x = F.relu(x)
    return conv1x1(x, params["mlp3_w"])


# --------------------------------------------------------------------------- #
# cal_sub_flow  (model.py:207-229)
# --------------------------------------------------------------------------- #
def cal_sub_flow(xyz, feature, interval, params, knn=16, return_stages=False, knn_fn=None):
    """xyz [B,3,5,h,w], feature [B,136,5,h,w], interval [B] ->
    flow [B,1,h,w], flow_prob [B,5,h,w].  ``knn_fn(xyz) -> idx`` replaces the
    canonical-order kNN (used by tests to replay the reference's tie order)."""
    B, _, M, h, w = xyz.shape
    # model.py:208 (kernel_size = 5 = len(interval_list))
    nn_idx = knn3d(xyz, M, knn) if knn_fn is None else knn_fn(xyz)
    x = feature.reshape(B, -1, M * h * w)
    outs = []
    for l in range(3):  # model.py:213-216
        x = edge_conv(x, nn_idx, params["ec%d_w1" % l], params["ec%d_w2" % l],
                      params["ec%d_gamma" % l], params["ec%d_beta" % l], concat_central=(l > 0))
        outs.append(x)
    cat = torch.cat(outs, dim=1)  # model.py:218
    raw = flow_mlp(cat, params).reshape(B, M, h, w)  # model.py:220-221
    prob = F.softmax(-raw, dim=1)  # model.py:222
    length = torch.tensor(HYPOTHESES).float().view(1, -1, 1, 1) * interval.view(-1, 1, 1, 1)
    flow = torch.sum(prob * length, dim=1, keepdim=True)  # model.py:224-227
    if return_stages:
        return flow, prob, {"nn_idx": nn_idx, "edge": outs, "raw": raw}
    return flow, prob


# --------------------------------------------------------------------------- #
# a2-a9: hypothesis points, multi-view fetch, variance  (model.py:150-204)
# --------------------------------------------------------------------------- #
def build_point_features(depth, interval, image_scale, pyramids, cam_params, mean, std,
                         img_hw, is_test=True, sub=None):
    """depth [B,1,hp,wp] (previous estimate), interval [B], pyramids = list of
    [B,V,C,hl,wl] (conv1, conv2, conv3), cam_params [B,V,2,4,4], mean/std [B,3],
    img_hw = (H, W) of the input images.
    Returns feature [B,136,5,h,w], xyz [B,3,5,h,w], depth_up [B,1,h,w].

    ``sub=(i, j, ratio)`` evaluates only the pixels (y*ratio+i, x*ratio+j) of ONE strided
    sub-cloud (model.py:236-255) -- same operations on a subset of the (independent) points, so
    that one sub-cloud of a large configuration can be checked without the full-size tensors;
    the returned tensors then have the sub-grid size (h/ratio, w/ratio)."""
    B, V = cam_params.shape[:2]
    H, W = img_hw
    ext = cam_params[:, :, 0, :3, :4]  # model.py:54
    R = ext[:, :, :, :3]
    t = ext[:, :, :, 3:4]
    R_inv = torch.inverse(R)  # model.py:57
    h, w = depth.shape[2:]
    if h != int(H * image_scale):  # model.py:153-158
        h, w = int(H * image_scale), int(W * image_scale)
        depth = F.interpolate(depth, (h, w), mode="nearest")
    K = cam_params[:, :, 1, :3, :3].clone()  # model.py:159-163
    K[:, :, :2, :3] *= image_scale if is_test else 4 * image_scale
    grid = get_pixel_grids(h, w).view(1, 1, 3, -1).expand(B, 1, 3, -1)
    uv = torch.matmul(torch.inverse(K[:, 0]).unsqueeze(1), grid)  # model.py:169-170
    oh, ow = h, w
    if sub is not None:
        si, sj, sr = sub
        sel = (torch.arange(si, h, sr).view(-1, 1) * w + torch.arange(sj, w, sr).view(1, -1)).reshape(-1)
        uv = uv[..., sel]
        depth = depth.reshape(B, 1, -1)[..., sel].view(B, 1, h // sr, w // sr)
        oh, ow = h // sr, w // sr
    feats, xyzs = [], []
    # model.py:180-185 resizes every level inside the hypothesis loop; the result does not depend
    # on the hypothesis, so it is computed once here (identical values)
    resized = []
    for level in pyramids:
        c, hl, wl = level.shape[2:]
        resized.append(F.interpolate(level.reshape(-1, c, hl, wl), (h, w), mode="bilinear",
                                     align_corners=False).view(B, V, c, h, w))
    for m in HYPOTHESES:  # model.py:173
        dm = depth + interval.view(-1, 1, 1, 1) * m
        cam_pts = uv * dm.view(B, 1, 1, -1)
        world = torch.matmul(R_inv[:, 0:1], cam_pts - t[:, 0:1]).transpose(1, 2).reshape(B, 3, -1)
        per_level = []
        for lv in resized:  # model.py:180-190
            pf = feature_fetch(lv, world, K, ext)
            avg = pf.mean(dim=1)
            avg2 = (pf ** 2).mean(dim=1)
            per_level.append(avg2 - avg ** 2)
        xyz = (world - mean.unsqueeze(-1)) / std.unsqueeze(-1)  # model.py:46-48,193
        per_level.append(xyz.repeat(1, 8, 1))  # model.py:194
        feats.append(torch.cat(per_level, dim=1))
        xyzs.append(xyz)
    feature = torch.stack(feats, dim=2).view(B, -1, len(HYPOTHESES), oh, ow)  # model.py:202
    xyz = torch.stack(xyzs, dim=2).view(B, 3, len(HYPOTHESES), oh, ow)  # model.py:203-204
    return feature, xyz, depth


# --------------------------------------------------------------------------- #
# (f-1) coarse cost volume: plane-sweep fetch + variance  (model.py:54-113)
# --------------------------------------------------------------------------- #
def coarse_cost_volume(feature_list, cam_params, is_test=True):
    """feature_list [B,V,C,h,w] (coarse_img_conv "conv3" per view, model.py:71-77),
    cam_params [B,V,2,4,4] -> cost volume [B,C,D,h,w] (model.py:113) and the depth
    hypotheses [B,D].  The reference view's fetched features are overwritten by the
    un-warped reference feature (model.py:103-106)."""
    B, V, C, h, w = feature_list.shape
    ext = cam_params[:, :, 0, :3, :4]
    R = ext[:, :, :, :3]
    t = ext[:, :, :, 3:4]
    R_inv = torch.inverse(R)
    K = cam_params[:, :, 1, :3, :3].clone()
    K[:, :, :2, :3] = K[:, :, :2, :3] / 2.0  # model.py:59
    if is_test:
        K[:, :, :2, :3] = K[:, :, :2, :3] / 4.0  # model.py:60-61
    depth_start = cam_params[:, 0, 1, 3, 0]
    depth_interval = cam_params[:, 0, 1, 3, 1]
    D = int(cam_params[0, 0, 1, 3, 2].long())
    depth_end = depth_start + (D - 1) * depth_interval  # model.py:67
    depths = torch.stack([torch.linspace(depth_start[i], depth_end[i], D) for i in range(B)], dim=0)  # [B,D]
    grid = get_pixel_grids(h, w).view(1, 1, 3, -1).expand(B, 1, 3, -1)
    uv = torch.matmul(torch.inverse(K[:, 0]).unsqueeze(1), grid)  # [B,1,3,hw]
    cam_pts = (uv.unsqueeze(3) * depths.view(B, 1, 1, D, 1)).view(B, 1, 3, -1)  # model.py:93
    world = torch.matmul(R_inv[:, 0:1], cam_pts - t[:, 0:1]).transpose(1, 2).contiguous().view(B, 3, -1)
    pf = feature_fetch(feature_list, world, K, ext)  # model.py:102
    ref = feature_list[:, 0].unsqueeze(2).expand(-1, -1, D, -1, -1).contiguous().view(B, C, -1)
    pf[:, 0] = ref  # model.py:103-106
    avg = pf.mean(dim=1)
    avg2 = (pf ** 2).mean(dim=1)
    return (avg2 - avg ** 2).view(B, C, D, h, w), depths


# --------------------------------------------------------------------------- #
# a1: one PointFlow iteration  (model.py:150-295)
# --------------------------------------------------------------------------- #
def point_flow(depth, interval, image_scale, pyramids, cam_params, mean, std, img_hw,
               params, is_test=True, knn=16, return_stages=False, knn_fn=None, sub=None):
    """Returns (flow_result [B,1,h,w], flow_prob [B,5,h,w]) for one iteration.
    ``sub=(i, j)`` (test mode, ratio > 1): only that strided sub-cloud is evaluated and the
    results have the sub-grid size -- one of the ratio^2 independent calls of model.py:236-267."""
    ratio = int(image_scale * 8) if is_test else 1
    if sub is not None:
        assert ratio > 1
        sub = (sub[0], sub[1], ratio)
    feature, xyz, depth_up = build_point_features(depth, interval, image_scale, pyramids,
                                                  cam_params, mean, std, img_hw, is_test, sub=sub)
    B, _, M, h, w = xyz.shape
    if ratio <= 1 or sub is not None:  # model.py:231-234 (test, scale 0.125) / :271-293 (train)
        flow, prob = cal_sub_flow(xyz, feature, interval, params, knn, knn_fn=knn_fn)
    else:  # model.py:236-267
        sh, sw = h // ratio, w // ratio
        f7 = feature.view(B, -1, M, sh, ratio, sw, ratio)
        x7 = xyz.view(B, 3, M, sh, ratio, sw, ratio)
        flow = torch.empty(B, 1, sh, ratio, sw, ratio)
        prob = torch.empty(B, M, sh, ratio, sw, ratio)
        for i in range(ratio):
            for j in range(ratio):
                fl, pr = cal_sub_flow(x7[:, :, :, :, i, :, j].contiguous(),
                                      f7[:, :, :, :, i, :, j].contiguous(), interval, params, knn,
                                      knn_fn=knn_fn)
                flow[:, :, :, i, :, j] = fl
                prob[:, :, :, i, :, j] = pr
        flow = flow.view(B, 1, h, w)
        prob = prob.view(B, M, h, w)
    result = depth_up + flow
    if return_stages:
        return result, prob, {"feature": feature, "xyz": xyz, "depth_up": depth_up, "flow": flow}
    return result, prob


def point_flow_pass(coarse_depth, depth_interval, pyramids, cam_params, mean, std, img_hw,
                    params, img_scales=(0.125, 0.25, 0.5), inter_scales=(1.0, 0.75, 0.15),
                    is_test=True, knn=16, knn_fn=None):
    """The iteration loop, model.py:297-303.  Returns the list of per-iteration
    (depth, prob)."""
    outs = []
    depth = coarse_depth
    for s, isc in zip(img_scales, inter_scales):
        depth, prob = point_flow(depth, isc * depth_interval, s, pyramids, cam_params, mean, std,
                                 img_hw, params, is_test, knn, knn_fn=knn_fn)
        outs.append((depth, prob))
    return outs


def params_from_state_dict(sd, prefix=""):
    """Map reference state_dict keys to the oracle's flat names."""
    p = {}
    for l in range(3):
        base = "%sflow_edge_conv.%d." % (prefix, l)
        p["ec%d_w1" % l] = sd[base + "conv1.weight"].float()
        p["ec%d_w2" % l] = sd[base + "conv2.weight"].float()
        p["ec%d_gamma" % l] = sd[base + "bn.weight"].float()
        p["ec%d_beta" % l] = sd[base + "bn.bias"].float()
    for i in range(3):
        base = "%sflow_mlp.0.%d." % (prefix, i)
        p["mlp%d_w" % i] = sd[base + "conv.weight"].float()
        p["mlp%d_gamma" % i] = sd[base + "bn.weight"].float()
        p["mlp%d_beta" % i] = sd[base + "bn.bias"].float()
    p["mlp3_w"] = sd[prefix + "flow_mlp.1.weight"].float()
    return p
