"""Where do the instructions of a kernel go?  Reads `ncu -i X.ncu-rep --page source --csv --print-source sass` (one kernel,
captured with --import-source on) and prints executed warp instructions and stall samples per block of SASS lines, then
the opcode mix of the ranges given as a:b arguments.
    python tools/ncu_source_mix.py report.ncu-rep [a:b ...] [--dump a:b]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[k - 1][1][:150] if k else "")
    hdr, data = rows[k], rows[k + 1:]
    ix = {h: i for i, h in enumerate(hdr)}
    I, S = ix["Instructions Executed"], ix["# Samples"]
    tot = sum(int(r[I]) for r in data)
    ts = sum(int(r[S]) for r in data) or 1
    print("warp instructions %d, samples %d, SASS lines %d" % (tot, ts, len(data)))
    for a in range(0, len(data), 100):
        seg = data[a:a + 100]
        e, s = sum(int(r[I]) for r in seg), sum(int(r[S]) for r in seg)
        if e:
            print("%5d  %11d %5.1f%%   samples %6d %5.1f%%" % (a, e, 100 * e / tot, s, 100 * s / ts))
    args = sys.argv[2:]
    dump = None
    if "--dump" in args:
        dump = args[args.index("--dump") + 1]
        args = [x for x in args if x not in ("--dump", dump)]
    for rng in args:
        a, b = (int(x) for x in rng.split(":"))
        c, st = collections.Counter(), collections.Counter()
        for r in data[a:b]:
            op = r[1].strip().split()
            if op and op[0].startswith("@"):
                op = op[1:]
            name = op[0].split(".")[0] if op else "?"
            c[name] += int(r[I])
            st[name] += 1
        t = sum(c.values()) or 1
        print("range %d:%d  %d warp instructions" % (a, b, t))
        for name, v in c.most_common(22):
            print("   %-12s %11d %5.1f%%  static %d" % (name, v, 100 * v / t, st[name]))
    if dump:
        a, b = (int(x) for x in dump.split(":"))
        for i in range(a, b):
            r = data[i]
            print(i, r[1][:72].strip().ljust(72), r[I].rjust(9), r[S].rjust(5))


if __name__ == "__main__":
    main()
