#!/usr/bin/env python
"""Turn an ncu report (gpurun_out/*.ncu-rep, `ncu --set full`) into the per-kernel CSV summary committed under
profiles/, and list the Blackwell SASS mnemonics of the built library.  Runs in the build container (no GPU).

    python tests/profile_summary.py ncu  gpurun_out/prof.ncu-rep  profiles/r02/ncu_it3_r02.csv
    python tests/profile_summary.py sass pointmvsnet_b200/libpmvs_b200.so  profiles/r02/sass_r02.txt
"""
import csv
import re
import subprocess
import sys

METRICS = [
    ("duration_us", "gpu__time_duration.sum"),
    ("dram_read_MB", "dram__bytes_read.sum"),
    ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct_of_peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pipe_active_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("l1tex_throughput_pct", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
    ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("alu_pipe_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("warp_inst_M", "smsp__inst_executed.sum"),
    ("registers", "launch__registers_per_thread"),
    ("dyn_smem_B", "launch__shared_mem_per_block_dynamic"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("shared_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
]


def ncu_summary(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [m[0] for m in METRICS])
        for r in data:
            name = re.sub(r"\(.*", "", r[idx["Kernel Name"]].replace("(anonymous namespace)", "anon").replace("<unnamed>", "anon")).replace("void ", "").strip()
            vals = []
            for key, met in METRICS:
                v = r[idx[met]].replace(",", "") if met in idx else ""
                try:
                    x = float(v)
                    if key == "warp_inst_M":
                        x /= 1e6
                    if key.endswith("_MB") and units[idx[met]] == "byte":
                        x /= 1e6
                    if key.endswith("_MB") and units[idx[met]] == "Kbyte":
                        x /= 1e3
                    if key.endswith("_MB") and units[idx[met]] == "Gbyte":
                        x *= 1e3
                    if key == "duration_us" and units[idx[met]] == "ns":
                        x /= 1e3
                    if key == "duration_us" and units[idx[met]] == "ms":
                        x *= 1e3
                    v = ("%.3f" % x).rstrip("0").rstrip(".")
                except ValueError:
                    pass
                vals.append(v)
            w.writerow([name] + vals)
    print("wrote", out, len(data), "kernels")


def sass_summary(lib, out):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    want = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "FFMA2", "FADD2",
            "FMUL2", "HMMA", "DSETP", "VIMNMX3"]
    per = {}
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur.replace("(anonymous namespace)", "anon"))
            per.setdefault(cur, {})
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for wname in want:
                if op == wname:
                    per[cur][wname] = per[cur].get(wname, 0) + 1
    with open(out, "w") as f:
        f.write("# Blackwell-specific SASS mnemonics per kernel of %s (cuobjdump -sass; sm_100a)\n" % lib)
        f.write("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCATOMSWS = tcgen05.alloc/dealloc, UTCBAR = tcgen05.commit,\n")
        f.write("# UTMALDG = cp.async.bulk.tensor (TMA tensor load), UBLKCP = cp.async.bulk, LDGSTS = cp.async, SYNCS = mbarrier ops,\n")
        f.write("# FFMA2/FADD2/FMUL2 = packed fp32 pairs, DSETP = fp64 compare (kNN keys), HMMA = legacy mma.sync (must be 0)\n")
        tot = {}
        for k in sorted(per):
            if per[k]:
                f.write("%-90s %s\n" % (k[:90], "  ".join("%s=%d" % kv for kv in sorted(per[k].items()))))
                for a, b in per[k].items():
                    tot[a] = tot.get(a, 0) + b
        f.write("TOTAL  " + "  ".join("%s=%d" % kv for kv in sorted(tot.items())) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "ncu":
        ncu_summary(sys.argv[2], sys.argv[3])
    else:
        sass_summary(sys.argv[2], sys.argv[3])
