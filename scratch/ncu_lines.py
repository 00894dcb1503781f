"""Per CUDA source line instruction / sample shares of one kernel of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python scratch/ncu_lines.py <rep> <kernel regex> [top N]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
H = None
lines = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        H = r
        ii, sa = H.index("Instructions Executed"), H.index("# Samples")
    elif H is not None and len(r) > ii and r[0] not in ("", "Line No") and r[0].isdigit():
        try:
            lines.append((cur_file, int(r[0]), r[1].strip(), int(r[ii] or 0), int(r[sa] or 0)))
        except ValueError:
            pass
tot = sum(l[3] for l in lines); ts = sum(l[4] for l in lines)
print("total warp instructions %d, samples %d" % (tot, ts))
byfile = {}
for f, n, src, i, s_ in lines:
    a = byfile.setdefault(f, [0, 0]); a[0] += i; a[1] += s_
for f, (i, s_) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("  %-24s %5.1f%% inst %5.1f%% samples" % (f, 100 * i / tot, 100 * s_ / max(ts, 1)))
for f, n, src, i, s_ in sorted(lines, key=lambda l: -l[3])[:top]:
    print("%5.1f%% %5.1f%%  %s:%d  %s" % (100 * i / tot, 100 * s_ / max(ts, 1), f, n, src[:110]))
