"""Turn gpurun_out ncu artefacts into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py <launches.csv> <prof.ncu-rep> <tag> [planes per profiled launch, default 8]

writes profiles/<tag>_launches.json (per-kernel time shares of the bench command) and
profiles/<tag>_mft_ncu.json (+ profiles/mft_ncu_summary.json, which bench.py reads for
roofline.traffic)."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NPLANES = int(sys.argv[4]) if len(sys.argv) > 4 else 8      # planes per launch of the profiled target (scripts/ncu_target.py N)


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        name = r[ik].split("(")[0].replace("void ", "")
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    return {"total_ms": total, "kernels": [
        {"kernel": k, "launches": cnt[k], "ms": tot[k], "avg_ms": tot[k] / cnt[k], "share": tot[k] / total}
        for k in sorted(tot, key=tot.get, reverse=True)]}


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
            "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[w] = v * scale.get(units[i], 1)
                if units[i] not in scale and units[i]:
                    d[w + ".unit"] = units[i]
        res.append(d)
    return res


if __name__ == "__main__":
    lpath, ppath, tag = sys.argv[1:4]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    L = launches(lpath)
    json.dump(L, open(os.path.join(ROOT, "profiles", f"{tag}_launches.json"), "w"), indent=1)
    F = full(ppath)
    json.dump(F, open(os.path.join(ROOT, "profiles", f"{tag}_mft_ncu.json"), "w"), indent=1)
    czt = [d for d in F if "czt_stage" in d["kernel"]]
    if czt:
        # one stage-A and one stage-B launch of NPLANES planes each: DRAM bytes of both stages per plane
        tr = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in czt) / (len(czt) / 2)
        json.dump({"dram_bytes_per_launch_pair": tr, "planes_per_profiled_launch": NPLANES, "dram_bytes_per_plane": tr / NPLANES,
                   "source": f"profiles/{tag}_mft_ncu.json",
                   "note": f"ncu --set full on scripts/ncu_target.py ({NPLANES} planes 1001^2->1024^2 per launch), czt_stage_kernel A + B",
                   "fp64_pipe_active_pct": sum(d.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 0) for d in czt) / len(czt),
                   "l1tex_throughput_pct": sum(d.get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", 0) for d in czt) / len(czt)},
                  open(os.path.join(ROOT, "profiles", "czt_ncu_summary.json"), "w"), indent=1)
    mft = [d for d in F if "mft_folded" in d["kernel"]]
    if mft:
        tr = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in mft) / len(mft)
        json.dump({"dram_bytes_per_launch": tr, "planes_per_profiled_launch": NPLANES,
                   "dram_bytes_per_plane_stage": tr / NPLANES, "source": f"profiles/{tag}_mft_ncu.json",
                   "note": f"ncu --set full on scripts/ncu_target.py ({NPLANES} planes 1001^2->1024^2 per launch)",
                   "tensor_pipe_active_pct": sum(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) for d in mft) / len(mft)},
                  open(os.path.join(ROOT, "profiles", "mft_ncu_summary.json"), "w"), indent=1)
    for k in L["kernels"]:
        print(f"{k['share']*100:6.2f}%  {k['avg_ms']:9.4f} ms x{k['launches']:3d}  {k['kernel']}")
