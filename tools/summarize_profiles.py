"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/ (round-tagged).

  launches.csv        <- ncu --metrics gpu__time_duration.sum --clock-control none ... (per-launch device time)
  prof_gemm.ncu-rep   <- ncu --set full --clock-control none --import-source on -k regex:gemm_kernel ...
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
go = os.path.join(ROOT, "gpurun_out")


def launch_list():
    rows = list(csv.reader(open(os.path.join(go, "launches.csv"))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    unit = rows[hi + 1][hdr.index("Metric Unit")] if "Metric Unit" in hdr else "ns"
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= iv:
            continue
        name = r[ik].split("(")[0].replace("void ", "")
        v = float(r[iv].replace(",", ""))
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f"# ncu launch list ({tag}): ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv python bench.py"
             " --steps 1 --warmup 1 --no-cpu-baseline --no-e2e",
             "# (the first 1200 launches of the bench command: the engine's eager warm-up of the forward and the step program,"
             " i.e. two DDPM++ UNet evaluations at B=512, plus the start of the first purification)",
             f"# per-launch device time, cold-cache/serialised:",
             f"# compare SHARES, not absolutes. unit of column 3: {unit}", "",
             f"{'kernel':70s} {'launches':>8s} {'total':>14s} {'share':>7s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k[:70]:70s} {v[0]:8d} {v[1]:14.0f} {100 * v[1] / tot:6.1f}%")
    open(os.path.join(out, f"{tag}_launch_list_eval.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


def gemm_capture(rep="prof_gemm.ncu-rep"):
    raw = subprocess.run(["ncu", "-i", os.path.join(go, rep), "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    lines = [f"# ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 4 python tests/gpu_profile_eval.py 512 1  ({tag})",
             "# dominant kernel dp::gemm_kernel<BN, EPI, CG> (tcgen05 implicit GEMM; CG = 2: CTA pair, cta_group::2).",
             "# launches: the two res-blocks after the input conv of one DDPM++ evaluation at B=512: conv 128->128 @32x32 (+temb, stats,",
             "# bf16 out) and conv 128->128 @32x32 (+fp32 residual, 1/sqrt2, stats), twice -- all four on CTA pairs", ""]
    traffic = []
    for ri, r in enumerate(rows[2:]):
        lines.append(f"## launch {ri}")
        rd = wr = None
        for w in want:
            for i, h in enumerate(hdr):
                if h == w:
                    lines.append(f"  {h:66s} {r[i]:>18s} {units[i]}")
                    if h == "dram__bytes_read.sum":
                        rd = float(r[i]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(units[i], 1)
                    if h == "dram__bytes_write.sum":
                        wr = float(r[i]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(units[i], 1)
        if rd is not None and wr is not None:
            traffic.append(rd + wr)
    open(os.path.join(out, f"{tag}_gemm_ncu_full.txt"), "w").write("\n".join(lines) + "\n")
    if traffic:
        json.dump({"dram_bytes_per_launch": sum(traffic) / len(traffic), "launches": len(traffic),
                   "source": f"profiles/{tag}_gemm_ncu_full.txt (dram__bytes_read.sum + dram__bytes_write.sum)",
                   "algorithmic_bytes_per_launch_note": "B=512 conv 128->128 @32x32: temb/bf16-out launches 268 MB, fp32-residual launches 671 MB (A bf16 134 MB + residual 268 MB + out 268 MB)"},
                  open(os.path.join(out, "gemm_dram_bytes_per_launch.json"), "w"), indent=1)
    print("\n".join(lines[:24]))
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "diffpure_b200", "libdiffpure_b200.so")],
                          capture_output=True, text=True).stdout
    cnt = collections.Counter()
    for m in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.4D.2CTA", "LDTM", "UTCBAR", "UTCBAR.2CTA.MULTICAST", "SYNCS.ARRIVE.TRANS64",
              "HMMA"):
        cnt[m] = sass.count(m)
    cnt["HMMA"] -= cnt["UTCHMMA"]   # "UTCHMMA" contains "HMMA": legacy mma.sync count is what is left
    open(os.path.join(out, f"{tag}_sass_mnemonics.txt"), "w").write(
        "# cuobjdump -sass diffpure_b200/libdiffpure_b200.so | mnemonic counts (tcgen05.mma -> UTCHMMA, TMA -> UTMALDG,\n"
        "# tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR; no legacy HMMA)\n" + "\n".join(f"{k:24s} {v}" for k, v in cnt.items()) + "\n")
    print(dict(cnt))


if __name__ == "__main__":
    launch_list()
    gemm_capture()
