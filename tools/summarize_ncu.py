"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the small, tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01b_ncu_launch_list_summary.csv "<header note>"
    python tools/summarize_ncu.py full gpurun_out/r01b_gemm_pair.ncu-rep profiles/r01b_ncu_gemm_pair.csv
"""
import csv
import subprocess
import sys
from collections import defaultdict

METRICS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_uniform.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second"]


def launches(src, dst, note):
    rows = list(csv.reader(l for l in open(src, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0][:96]
        tot[name] += float(r[iv].replace(",", ""))
        cnt[name] += 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# {note}\n# total {total / 1e6:.2f} ms over {sum(cnt.values())} launches; cold-cache serialised times: compare SHARES\n")
        f.write("share_pct,total_ns,launches,kernel\n")
        for k in sorted(tot, key=lambda k: -tot[k]):
            f.write(f"{100 * tot[k] / total:.2f},{tot[k]:.0f},{cnt[k]},{k}\n")


def launches_dram(src, dst, note):
    """launch list captured with gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum: per kernel name the time
    share, launches, DRAM bytes per launch and DRAM GB/s."""
    rows = list(csv.reader(l for l in open(src, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    ii, ik, iv, im = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    per = defaultdict(lambda: defaultdict(float))
    name_of = {}
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        per[r[ii]][r[im]] = float(r[iv].replace(",", ""))
        name_of[r[ii]] = r[ik].split("(")[0][:96]
    tot, cnt, byt = defaultdict(float), defaultdict(int), defaultdict(float)
    for i, m in per.items():
        n = name_of[i]
        tot[n] += m.get("gpu__time_duration.sum", 0.0)
        byt[n] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        cnt[n] += 1
    total = sum(tot.values())
    lib = [k for k in tot if "ged::" not in k and not k.startswith("ged")]
    with open(dst, "w") as f:
        f.write(f"# {note}\n# total {total / 1e6:.2f} ms over {sum(cnt.values())} launches (cold-cache, serialised: compare SHARES); "
                f"library (non ged::) kernels: {sum(cnt[k] for k in lib)} launches = {100 * sum(tot[k] for k in lib) / total:.2f} % of the time\n")
        f.write("share_pct,total_ms,launches,dram_MB_per_launch,dram_GBs,kernel\n")
        for k in sorted(tot, key=lambda k: -tot[k]):
            f.write(f"{100 * tot[k] / total:.2f},{tot[k] / 1e6:.3f},{cnt[k]},{byt[k] / cnt[k] / 1e6:.2f},{byt[k] / max(tot[k], 1):.0f},{k}\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = [hdr.index("Kernel Name")] + [hdr.index(m) for m in METRICS if m in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in keep])
        w.writerow([units[i] for i in keep])
        for r in rows[2:]:
            w.writerow([r[i][:110] for i in keep])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    elif sys.argv[1] == "launches_dram":
        launches_dram(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
