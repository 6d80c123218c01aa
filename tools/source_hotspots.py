"""Source-level hot spots of a kernel from an `ncu --set full --import-source on` capture.

ncu's CLI exports the SASS view with per-instruction counters (`--page source --csv`) but without the CUDA line of each
instruction; `nvdisasm -g` of the object that was profiled prints the same SASS with `//## File ..., line N` markers.
The two listings are the same instruction sequence, so they are joined by position (the script refuses to join when the
instruction counts differ, i.e. when the object was rebuilt with different code for that kernel since the capture).

    python tools/source_hotspots.py gpurun_out/prof_r01j.ncu-rep structure-plp-slam_b200/build/match.o point_match_kernel [top]
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def disasm(obj: Path, kernel: str):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(obj.resolve())], cwd=td, capture_output=True, check=True)
        cubins = list(Path(td).glob("*.cubin"))
        out = subprocess.run(["nvdisasm", "-g", "-c", str(cubins[0])], capture_output=True, text=True, check=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kernel in l)
    end = next((i for i, l in enumerate(lines) if i > start and l.startswith("//---------------------")), len(lines))
    cur, seq = None, []
    for l in lines[start:end]:
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            seq.append(cur)
    return seq


def ncu_sass(rep: Path, kernel: str):
    """-> (header, [rows of launch 0, rows of launch 1, ...]) -- one section per captured launch of a matching kernel"""
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    header, sections = None, []
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "Kernel Name":
            sections.append([])
        elif r[0] == "Address":
            header = r
        elif r[0].startswith("0x") and sections:
            sections[-1].append(r)
    return header, sections


def main():
    rep, obj, kernel = Path(sys.argv[1]), Path(sys.argv[2]), sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    # "name" or "name:mangled-fragment" (template instances: lsd_grow_kernel:lsd_grow_kernelILb1E)
    kernel, _, sym = kernel.partition(":")
    seq = disasm(obj, sym or kernel)
    h, sections = ncu_sass(rep, kernel)
    ix = {n: i for i, n in enumerate(h)}
    sections = [sec for sec in sections if len(sec) == len(seq)]  # other template instances / rebuilt code drop out
    if not sections:
        sys.exit(f"{kernel}: no captured launch has the {len(seq)} instructions of {obj.name} -- object rebuilt since?")
    launches = len(sections)
    stall, inst = collections.Counter(), collections.Counter()
    for sec in sections:
        for cur, r in zip(seq, sec):
            stall[cur] += float(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
            inst[cur] += float(r[ix["Instructions Executed"]] or 0)
    ts, ti = sum(stall.values()) or 1.0, sum(inst.values()) or 1.0
    srcs = {}
    print(f"### {kernel} ({launches} captured launch(es), {ti / launches / 1e6:.2f} M warp instructions per launch)\n")
    print("| source line | stall samples % | instructions % | text |")
    print("|---|---|---|---|")
    for cur, v in sorted(stall.items(), key=lambda kv: -kv[1])[:top]:
        f, l = cur if cur else ("?", 0)
        if f not in srcs:
            c = list((ROOT / "structure-plp-slam_b200" / "csrc").glob(f))
            srcs[f] = c[0].read_text().splitlines() if c else []
        text = srcs[f][l - 1].strip()[:88].replace("|", "\\|") if srcs[f] and 0 < l <= len(srcs[f]) else ""
        print(f"| {f}:{l} | {100 * v / ts:.1f} | {100 * inst[cur] / ti:.1f} | `{text}` |")
    print()


if __name__ == "__main__":
    main()
