"""Annotated SASS of the hottest loop of a capture: python scripts/sass_excerpt.py <prefix of .source.csv> <out file>
(source page exported by scripts/profile_r02.sh; developer tool)."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1] + ".source.csv")))
    kern = ["", ""]
    while rows and not (len(rows[0]) > 2 and rows[0][0].strip().lower() == "address"):
        kern = rows.pop(0)
    hdr, body = rows[0], rows[1:]
    ci = {h.strip().lower(): k for k, h in enumerate(hdr)}
    A, S, W, E = ci["address"], ci["source"], ci["warp stall sampling (all samples)"], ci["instructions executed"]
    smp = [float(r[W] or 0) for r in body]
    tot = sum(smp)
    addr = [int(r[A], 16) for r in body]
    pos = {a: k for k, a in enumerate(addr)}
    # every backward branch closes a loop; take the innermost loop (fewest instructions) with 4 gathers that holds most samples
    best = None
    for k, r in enumerate(body):
        s = r[S]
        if "BRA" in s and "0x" in s:
            try:
                tgt = int(s.split("0x")[-1].strip().rstrip(";"), 16)
            except ValueError:
                continue
            if tgt in pos and pos[tgt] < k:
                lo = pos[tgt]
                text = " ".join(body[i][S] for i in range(lo, k + 1))
                if text.count("LDG.E.64") == 4 and k - lo < 140:
                    share = sum(smp[lo:k + 1])
                    if best is None or share > best[0]:
                        best = (share, lo, k)
    share, lo, hi = best
    out = ["# SASS of the hottest loop of %s (Gram of 192 x 500-tip trees, fp64): the two-gather row loop, two rows in flight." % kern[1],
           "# From the source page of the ncu capture (%s): offset in the function, share of all warp-stall samples, warp instructions executed, SASS." % sys.argv[1],
           "# %d instructions per two rows of 32 cells; the loop holds %.1f %% of the kernel's stall samples." % (hi - lo + 1, 100 * share / tot)]
    for k in range(lo, hi + 1):
        out.append("%06x  %5.2f%%  %9s  %s" % (addr[k] - addr[0], 100 * smp[k] / tot, body[k][E], body[k][S].strip()))
    open(sys.argv[2], "w").write("\n".join(out) + "\n")
    print("\n".join(out[:3]))


if __name__ == "__main__":
    main()
