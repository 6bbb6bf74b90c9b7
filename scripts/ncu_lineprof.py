#!/usr/bin/env python
"""Per-source-line and per-region view of an ncu capture of score_stream_kernel.

`ncu --page source --csv` lists SASS instructions without line numbers; this
joins them, in order, with `nvdisasm --print-line-info` of the same build.

    ncu -i gpurun_out/x.ncu-rep --page source --csv > /tmp/x_source.csv
    python scripts/ncu_lineprof.py /tmp/x_source.csv [top_lines]

The object file must be the build that was profiled
(nxsearch_b200/lib/obj/engine.cu.o).  The kernel instance is taken from the
first line of the csv.  Used for profiles/r1_v6_stream_regions.txt.
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "nxsearch_b200" / "csrc" / "gpu" / "stream.cuh"

MANGLED = {  # score_stream_kernel<LOGIC, WIDE, ALGO>
    (0, 0, 1): "_Z19score_stream_kernelILb0ELb0ELi1EEv12StreamParams",
    (0, 0, 0): "_Z19score_stream_kernelILb0ELb0ELi0EEv12StreamParams",
    (1, 0, 0): "_Z19score_stream_kernelILb1ELb0ELi0EEv12StreamParams",
    (1, 0, 1): "_Z19score_stream_kernelILb1ELb0ELi1EEv12StreamParams",
}


def disassemble(symbol: str):
    """[(line-info, text)] of one function, in address order."""
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "nxsearch_b200/lib/obj/engine.cu.o")],
                       cwd=tmp, check=True, capture_output=True)
        cubin = next(Path(tmp).glob("*.cubin"))
        text = subprocess.run(["nvdisasm", "--print-line-info", str(cubin)], check=True,
                              capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(text) if l.startswith(".text." + symbol + ":"))
    cur, out = None, []
    for l in text[start + 1:]:
        if l.startswith("//-----") or l.lstrip().startswith(".section"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def regions(src):
    """Line ranges of the kernel's phases, found by their marker comments."""
    def find(txt, start=0):
        return next(i + 1 for i, l in enumerate(src[start:], start) if txt in l)
    prod, cons = find("================= producer warp"), find("================= consumer warps")
    dense, full = find("if (!WIDE && (flags & ST_F_DENSE))"), find("} else if (flags & ST_F_FULL)")
    part = find("const uint32_t nsub = (flags >> ST_F_NSUB_SHIFT) & 0x3fu;", full)
    rel = find("Stage consumed.  A sparse item")
    epi = find("item epilogue: top-k of the tile")
    pushed = find("if (ths != __uint_as_float(0x7f800000u) && npush <= PUSH)")
    sparse = find("} else if (flags & ST_F_SPARSE) {", pushed)
    scan = find("const float ths = thr_key ? __uint_as_float((uint32_t)(thr_key >> 32))", sparse + 8)
    rank = find("The other parity's counter was last read an item ago")
    end = find("finalize_cells_kernel")
    kern = find("score_stream_kernel(const StreamParams p)")
    return [(kern, prod - 1, "prologue"), (prod, cons - 1, "producer"), (cons, dense - 1, "consumer: wait / first stage / zero-fill"),
            (dense, full - 1, "dense run"), (full, part - 1, "full stage"), (part, rel - 2, "partial stage"),
            (rel - 1, epi - 1, "stage release"), (epi, pushed - 1, "epilogue: entry"), (pushed, sparse - 1, "epilogue: noted documents"),
            (sparse, scan - 3, "epilogue: sparse collect"), (scan - 2, rank - 1, "epilogue: scan + zero"),
            (rank, end, "epilogue: rank + emit")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    m = re.search(r"score_stream_kernel<\(bool\)(\d), \(bool\)(\d), \(int\)(\d)>", rows[0][1])
    ins = disassemble(MANGLED[tuple(int(x) for x in m.groups())])
    hdr, data = rows[1], rows[2:]
    col = {n: i for i, n in enumerate(hdr)}
    assert len(data) == len(ins), f"{len(data)} profiled instructions vs {len(ins)} disassembled: different build?"
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    src = SRC.read_text().split("\n")
    regs = regions(src)

    def region_of(n):
        return next((name for a, b, name in regs if a <= n <= b), None)

    by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    by_reg = collections.defaultdict(lambda: [0, 0])
    stall_tot = collections.Counter()
    tot_i = tot_s = 0
    cur = "prologue"
    for (ln, _), r in zip(ins, data):
        i, s = int(r[col["Instructions Executed"]]), int(r[col["# Samples"]])
        if ln and ln[0] == "stream.cuh" and region_of(ln[1]):
            cur = region_of(ln[1])        # inlined helpers count for the phase that calls them
        by_line[ln][0] += i
        by_line[ln][1] += s
        by_reg[cur][0] += i
        by_reg[cur][1] += s
        tot_i += i
        tot_s += s
        for c in stall_cols:
            v = int(r[col[c]] or 0)
            if v:
                by_line[ln][2][c[6:]] += v
                stall_tot[c[6:]] += v
    print(f"{rows[0][1]}\nwarp instructions {tot_i}, samples {tot_s}")
    print("stall reasons, % of samples:", ", ".join(f"{k} {100 * v / tot_s:.1f}" for k, v in stall_tot.most_common(10)))
    print("\nby phase (share of warp instructions / of samples):")
    for k, (i, s) in sorted(by_reg.items(), key=lambda x: -x[1][1]):
        print(f"  {k:42s} {100 * i / tot_i:5.1f}%  {100 * s / tot_s:5.1f}%")
    print(f"\ntop {top} source lines by samples:")
    for ln, (i, s, st) in sorted(by_line.items(), key=lambda x: -x[1][1])[:top]:
        f, n = ln if ln else ("?", 0)
        text = src[n - 1].strip()[:72] if f == "stream.cuh" and 0 < n <= len(src) else ""
        reasons = ",".join(f"{k}:{v}" for k, v in st.most_common(3))
        print(f"  {f}:{n:<5d} inst {100 * i / tot_i:4.1f}%  samples {100 * s / tot_s:4.1f}%  [{reasons}]  {text}")


if __name__ == "__main__":
    main()
