"""Summarise a BSMR_TRACE=all log of the stage clustering kernel (debug library): per range of stages the spacing of the
stage starts (= the founding chain, the critical path of the pipeline), the founding time and the stage duration.

    python profiles/stage_trace_summary.py gpurun_out/stage_trace20.log [ranges=12]
"""
import re
import sys


def main():
    path = sys.argv[1]
    nr = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    rows = []
    head = []
    for line in open(path, errors="replace"):
        m = re.match(r"\[bsmr trace\]\s+(\d+):\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)", line)
        if m:
            rows.append((int(m.group(1)), float(m.group(2)), float(m.group(3)), float(m.group(4))))
        elif line.startswith("[bsmr trace]") or line.startswith("{"):
            head.append(line.strip()[:1200])
    for h in head:
        print(h)
    if not rows:
        return
    rows.sort()
    n = len(rows)
    print("stages %d, last start %.1f ms, last end %.1f ms" % (n, rows[-1][1] / 1e3, rows[-1][3] / 1e3))
    print("%14s %12s %14s %14s %14s" % ("stages", "start ms", "spacing us", "founding us", "duration us"))
    step = max(1, n // nr)
    for a in range(0, n, step):
        b = min(n, a + step)
        seg = rows[a:b]
        spacing = (seg[-1][1] - seg[0][1]) / max(1, len(seg) - 1)
        founding = [r[2] - r[1] for r in seg if 0 < r[2] < 1e12]
        dur = [r[3] - r[1] for r in seg if r[3] > 0]
        print("%6d..%-6d %12.1f %14.1f %14.1f %14.1f" % (seg[0][0], seg[-1][0], seg[0][1] / 1e3, spacing,
                                                      sum(founding) / max(1, len(founding)), sum(dur) / max(1, len(dur))))


if __name__ == "__main__":
    main()
