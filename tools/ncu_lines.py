"""Per-source-line view of one kernel launch in an ncu report: where the warp instructions and the stall samples go.

    python tools/ncu_lines.py REPORT.ncu-rep LAUNCH_INDEX KERNEL_SUBSTRING [LIB.so] [--top N] [--sass N]

`ncu --page source --csv` exports counters per SASS instruction only; this joins them, in instruction order, with the line
table of the same kernel in the library (`cuobjdump -xelf` + `nvdisasm -g`; the library must be the build that was profiled:
the instruction counts are checked) and sums per line of the .cu file. The report must have been taken with
`--set full --import-source on` and the library built with -lineinfo. Used for profiles/ncu_r02_*_lines.txt.
"""
import csv
import os
import pickle
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_rows(report, launch):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    name = rows[0][1]
    hdr = rows[1]
    data, seen = [], set()
    for r in rows[2:]:
        if len(r) < len(hdr) - 2 or r[0] in seen:       # the export lists every instruction once per source view
            continue
        seen.add(r[0])
        data.append(r)
    return name, hdr, data


def line_table(lib, kernel):
    """[(offset, sass text, (file, line))] of the first kernel of `lib` whose mangled name contains `kernel`."""
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
        for f in sorted(os.listdir(d)):
            if not f.endswith(".cubin"):
                continue
            dis = subprocess.run(["nvdisasm", "-g", os.path.join(d, f)], capture_output=True, text=True).stdout
            if kernel not in dis:
                continue
            lines = dis.split("\n")
            start = end = None
            for i, l in enumerate(lines):
                if l.startswith(".text.") and kernel in l and start is None:
                    start = i
                elif l.startswith(".text.") and start is not None and i > start:
                    end = i
                    break
            if start is None:
                continue
            cur, ins = None, []
            for l in lines[start:end]:
                m = re.search(r'//## File "([^"]+)", line (\d+)', l)
                if m:
                    cur = (m.group(1), int(m.group(2)))
                    continue
                m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
                if m:
                    ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
            return ins
    raise SystemExit(f"no kernel containing {kernel!r} in {lib}")


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    nsass = int(sys.argv[sys.argv.index("--sass") + 1]) if "--sass" in sys.argv else 12
    args = [a for a in args if not a.isdigit() or a == args[1]]
    report, launch, kernel = args[0], int(args[1]), args[2]
    lib = args[3] if len(args) > 3 else os.path.join(ROOT, "smallk_b200", "lib", "libsmallk_b200.so")
    name, hdr, data = sass_rows(report, launch)
    ix = {h: i for i, h in enumerate(hdr)}

    def g(r, k):
        try:
            return int(r[ix[k]])
        except Exception:
            return 0
    tot_i = sum(g(r, "Instructions Executed") for r in data)
    tot_s = sum(g(r, "# Samples") for r in data)
    print(f"# {name[:160]}")
    print(f"# launch {launch} of {report}: {tot_i} warp instructions, {tot_s} stall samples, {len(data)} SASS instructions")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("# stall samples: " + ", ".join(f"{h[6:]} {100.0 * sum(g(r, h) for r in data) / max(tot_s, 1):.1f}%" for h in
                                          sorted(stalls, key=lambda h: -sum(g(r, h) for r in data))[:8]))
    mix = Counter()
    for r in data:
        t = r[ix["Source"]].split()
        op = t[1] if t[0].startswith("@") else t[0]
        mix[op.split(".")[0]] += g(r, "Instructions Executed")
    print("# instruction mix: " + ", ".join(f"{k} {v / 1e6:.1f}M" for k, v in mix.most_common(14)))
    print(f"\n## top SASS instructions by stall samples")
    for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:nsass]:
        why = max(stalls, key=lambda h: g(r, h))
        print(f"  {g(r, '# Samples'):8d} samples {g(r, 'Instructions Executed'):11d} exec  {r[ix['Source']].strip()[:64]:64s} ({why[6:]})")
    ins = line_table(lib, kernel)
    if 0 < abs(len(ins) - len(data)) <= 2:
        # nvdisasm and the report disagree about a trailing padding instruction or two: the common prefix is the same code
        keep = min(len(ins), len(data))
        ins, data = ins[:keep], data[:keep]
    if len(ins) != len(data):
        print(f"\n(the library's kernel has {len(ins)} SASS instructions, the report {len(data)}: not the profiled build, no per-line table)")
        return
    agg = defaultdict(lambda: [0, 0, 0])
    for r, (_, _, cur) in zip(data, ins):
        a = agg[cur]
        a[0] += g(r, "Instructions Executed"); a[1] += g(r, "# Samples"); a[2] += 1
    cache = {}

    def text(cur):
        if not cur:
            return ""
        f, l = cur
        if f not in cache:
            try:
                cache[f] = open(f).read().split("\n")
            except OSError:
                cache[f] = None
        src = cache[f]
        return src[l - 1].strip()[:104] if src and 0 < l <= len(src) else "(header)"
    print(f"\n## per source line (sorted by warp instructions)\n  {'file:line':30s} inst%  samples%   SASS  source")
    for cur, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        where = f"{os.path.basename(cur[0])}:{cur[1]}" if cur else "?"
        print(f"  {where:30s} {100 * a[0] / max(tot_i, 1):5.1f}  {100 * a[1] / max(tot_s, 1):7.1f}  {a[2]:6d}  {text(cur)}")


if __name__ == "__main__":
    main()
