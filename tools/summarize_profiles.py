"""Summarise gpurun_out/{launches_TAG.csv, prof_TAG.ncu-rep, bench_default.json} into profiles/ (tracked).

Run in the authoring container after a `gpurun -- bash tools/profile.sh TAG BATCH` call:
    python tools/summarize_profiles.py r01 64
"""
import collections
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
           "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
           "l1tex__t_bytes.sum", "lts__t_bytes.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__cycles_active.avg"]


def short(name):
    return name.split("(")[0].split("::")[-1]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    PROF.mkdir(exist_ok=True)
    lines = [f"# ncu summary {tag} (bench.py --batch {batch}; B200, clocks not locked)", ""]
    # ---- launch list
    lf = OUT / f"launches_{tag}.csv"
    if lf.exists():
        shutil.copy(lf, PROF / f"launches_{tag}.csv")
        rows = list(csv.reader(open(lf)))
        h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        hdr, data = rows[h], rows[h + 1:]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = collections.OrderedDict()
        for r in data:
            if len(r) <= vi:
                continue
            v = float(r[vi].replace(",", ""))
            v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
            agg.setdefault(short(r[ki]), []).append(v)
        tot = sum(sum(v) for v in agg.values())
        lines += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, 2 steps; cold-cache, "
                  "serialised: compare SHARES)", "", "| kernel | launches | total us | share | mean us |", "|---|---|---|---|---|"]
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            lines.append(f"| {k} | {len(v)} | {sum(v):.1f} | {sum(v) / tot:.3f} | {sum(v) / len(v):.1f} |")
        lines.append("")
    # ---- full capture
    rep = OUT / f"prof_{tag}.ncu-rep"
    if rep.exists():
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {m: hdr.index(m) for m in METRICS if m in hdr}
        kn = hdr.index("Kernel Name")
        with open(PROF / f"prof_{tag}_raw_selected.csv", "w", newline="") as fh:  # csv.writer quotes "kernel<1024, 3840, 2>"
            wr = csv.writer(fh)
            wr.writerow(["kernel"] + list(idx))
            for r in rows[2:]:
                wr.writerow([short(r[kn])] + [r[i] for i in idx.values()])
        lines += ["## `ncu --set full` capture of one bench step (one launch per kernel)", "",
                  "| kernel | grid x block | ms | DRAM read MB | DRAM write MB | DRAM % of peak | SM % of peak | warps active % | regs | waves/SM | inst (M) |",
                  "|---|---|---|---|---|---|---|---|---|---|---|"]
        for r in rows[2:]:
            g = lambda m: r[idx[m]] if m in idx else "n/a"
            lines.append(f"| {short(r[kn])} | {g('launch__grid_size')} x {g('launch__block_size')} | "
                         f"{float(g('gpu__time_duration.sum')):.4f} | {float(g('dram__bytes_read.sum')):.2f} | "
                         f"{float(g('dram__bytes_write.sum')):.2f} | {float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.2f} | "
                         f"{float(g('sm__throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
                         f"{float(g('sm__warps_active.avg.pct_of_peak_sustained_active')):.1f} | {g('launch__registers_per_thread')} | "
                         f"{g('launch__waves_per_multiprocessor')} | {float(g('smsp__inst_executed.sum')) / 1e6:.1f} |")
        lines.append("")
    bj = OUT / "bench_default.json"
    if bj.exists():
        txt = [l for l in bj.read_text().splitlines() if l.startswith("{")]
        if txt:
            d = json.loads(txt[-1])
            (PROF / f"bench_{tag}.json").write_text(json.dumps(d, indent=1) + "\n")
            lines += [f"## bench.py default run ({tag})", "",
                      f"* value = {d['value']:.0f} {d['unit']} (device-resident), e2e = {d['e2e']['value']:.0f} {d['unit']}, "
                      f"ms/step = {d['ms_per_step']:.2f}, launches in timed region = {d['gpu_launches']}",
                      f"* clocks: {d['clocks']}",
                      f"* roofline: {json.dumps({k: v for k, v in d['roofline'].items() if k != 'kernel_time_shares'})}",
                      f"* kernel time shares (CUDA events): {d['roofline']['kernel_time_shares']}",
                      f"* cpu_baseline: {d.get('cpu_baseline')}", ""]
    (PROF / f"summary_{tag}.md").write_text("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
