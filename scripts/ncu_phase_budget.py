"""Per-phase instruction budget of one kernel launch from an `ncu --set full --import-source on` capture.

ncu's source page lists, per SASS instruction, the warp-level `Instructions Executed`, the thread-level count and the stall
samples; `nvdisasm -g` of the same cubin gives the (file, line) every SASS instruction was compiled from (built with
-lineinfo). This script joins the two by instruction order and sums them per source line and per PHASE, where a phase is a
set of (file, first line, last line) ranges given in a small JSON file:

    {"node step": [["wf_trace8.cuh", 49, 96]], "triangle test": [["wf_intersect.cuh", 50, 74]], ...}

Lines that match no range go to "other". `nvdisasm -gi` gives the whole inlining chain of every instruction; an instruction belongs
to the phase of the innermost frame that some range lists (so helpers shared by two phases are simply left out of the ranges).

usage: python scripts/ncu_phase_budget.py <file.ncu-rep> <kernel regex> <launch index among matches> <mangled-name substring> <phases.json> [units per launch]
Writes a markdown table to stdout. Needs no GPU.
"""
import csv
import io
import json
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def sass_rows(rep, regex, index):
    """Per-instruction rows of the index-th (1-based) captured launch whose kernel name matches `regex`."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    # (ncu prints every launch's table twice in this mode: keep one of two identical consecutive sections)
    uniq = []
    for k, i in enumerate(starts):
        j = starts[k + 1] if k + 1 < len(starts) else len(rows)
        if uniq and rows[uniq[-1][0]:uniq[-1][1]] == rows[i:j]:
            continue
        uniq.append((i, j))
    starts = [i for i, _ in uniq]
    match = [i for i in starts if re.search(regex, rows[i][1])]
    if int(index) > len(match):
        raise SystemExit(f"only {len(match)} launches match /{regex}/")
    s0 = match[int(index) - 1]
    s1 = dict(uniq)[s0]
    name = rows[s0][1]
    hdr = rows[s0 + 1]
    col = {k: hdr.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples")}
    res = []
    for r in rows[s0 + 2:s1]:
        if len(r) < len(hdr):
            continue
        res.append((r[col["Source"]].strip(), int(r[col["Instructions Executed"]]), int(r[col["Thread Instructions Executed"]]), int(r[col["# Samples"]])))
    return name, res


def line_table(so_path, symbol_part):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(so_path)], cwd=td, capture_output=True)
        for cubin in sorted(Path(td).glob("*.cubin")):
            txt = subprocess.run(["nvdisasm", "-gi", "-c", str(cubin)], capture_output=True, text=True).stdout
            m = re.search(r"^\.text\.(\S*" + re.escape(symbol_part) + r"\S*):\n", txt, re.M)
            if not m:
                continue
            body = txt[m.end():]
            end = re.search(r"^//-+ \.", body, re.M)
            body = body[:end.start()] if end else body
            # -gi prints, before an instruction, the chain of inlined frames (innermost first); it stays valid until the next chain
            chain, fresh = [("?", 0)], True
            lines = []
            for ln in body.splitlines():
                f = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
                if f:
                    if fresh:
                        chain, fresh = [], False
                    chain.append((Path(f.group(1)).name, int(f.group(2))))
                    continue
                if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
                    fresh = True
                    lines.append((list(chain), ln.split("*/", 1)[1].strip()))
            return m.group(1), lines
    raise SystemExit(f"no function matching {symbol_part!r} in {so_path}")


def main():
    rep, regex, index, sym, phases_file = sys.argv[1:6]
    units = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
    phases = json.loads(Path(phases_file).read_text())
    name, rows = sass_rows(rep, regex, index)
    fn, lines = line_table(ROOT / "xraytracer_b200" / "csrc" / "libxrtgpu.so", sym)
    if len(rows) != len(lines):
        print(f"warning: {len(rows)} instructions in the capture, {len(lines)} in the library (rebuilt since the capture?)", file=sys.stderr)
    n = min(len(rows), len(lines))

    def phase_of(chain):
        # innermost frame first: code inlined from a shared helper belongs to the phase of the nearest caller that is listed
        for file, line in chain:
            for label, ranges in phases.items():
                for f, lo, hi in ranges:
                    if file == f and lo <= line <= hi:
                        return label
        return "other"

    agg = {}
    for k in range(n):
        chain, _ = lines[k]
        _, inst, tinst, samples = rows[k]
        a = agg.setdefault(phase_of(chain), [0, 0, 0, 0])
        a[0] += inst; a[1] += tinst; a[2] += samples; a[3] += 1
    tot = [sum(a[i] for a in agg.values()) for i in range(4)]
    print(f"kernel `{name.split('(')[0]}` (launch {index} of /{regex}/ in {Path(rep).name}): {tot[0] / 1e6:.1f} M warp instructions, "
          f"{tot[1] / max(tot[0], 1):.1f} active lanes on average, {tot[3]} SASS instructions\n")
    print("| phase | SASS instr | warp instr (M) | share | active lanes | stall samples | share |" + (" warp instr per unit |" if units else ""))
    print("|---|---|---|---|---|---|---|" + ("---|" if units else ""))
    for label in list(phases) + ["other"]:
        if label not in agg:
            continue
        a = agg[label]
        row = f"| {label} | {a[3]} | {a[0] / 1e6:.1f} | {100 * a[0] / max(tot[0], 1):.1f} % | {a[1] / max(a[0], 1):.1f} | {a[2]} | {100 * a[2] / max(tot[2], 1):.1f} % |"
        if units:
            row += f" {a[0] / units:.1f} |"
        print(row)


if __name__ == "__main__":
    main()
