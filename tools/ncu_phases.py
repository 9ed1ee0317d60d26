"""Per-phase summary of one kernel from an ncu report captured with --import-source on (run here, on the CPU box):
the SASS source page is cut into segments at BAR.SYNC; per segment: warp-instructions per CTA, sampled stall
reasons, opcode mix, shared-memory wavefronts.
  python tools/ncu_phases.py gpurun_out/r2_full.ncu-rep k1_transform [ctas]
"""
import collections
import csv
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1])
    hdr = rows[1]
    col = {k: i for i, k in enumerate(hdr)}
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    segs = [dict(inst=0, samples=0, stall=collections.Counter(), ops=collections.Counter(), wave=0, wave_ideal=0)]
    for r in rows[2:]:
        if len(r) < len(hdr) or r[0] == "Address":
            continue
        if r[0] == "Kernel Name":      # a second kernel instance of the same name follows: the first one is enough
            break
        sass = r[col["Source"]].strip()
        op = sass.split()[0] if sass else ""
        if op.startswith("@"):
            op = sass.split()[1]
        op = op.split(".")[0]
        s = segs[-1]
        n = int(r[col["Instructions Executed"]] or 0)
        s["inst"] += n
        s["ops"][op] += n
        s["samples"] += int(r[col["# Samples"]] or 0)
        s["wave"] += int(r[col["L1 Wavefronts Shared"]] or 0)
        s["wave_ideal"] += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
        for k in stalls:
            v = int(r[col[k]] or 0)
            if v:
                s["stall"][k[6:]] += v
        if sass.startswith("BAR.SYNC"):
            segs.append(dict(inst=0, samples=0, stall=collections.Counter(), ops=collections.Counter(), wave=0, wave_ideal=0))
    ctas = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    tot_i, tot_s = sum(s["inst"] for s in segs), sum(s["samples"] for s in segs)
    print(f"warp-instructions per CTA {tot_i / ctas:.0f}, samples {tot_s}")
    for i, s in enumerate(segs):
        if not s["inst"]:
            continue
        print(f"segment {i}: inst/CTA {s['inst'] / ctas:.0f} ({100 * s['inst'] / tot_i:.1f} %), samples {s['samples']} "
              f"({100 * s['samples'] / max(tot_s, 1):.1f} %), smem wavefronts/CTA {s['wave'] / ctas:.0f} (ideal {s['wave_ideal'] / ctas:.0f})")
        print("   stalls:", ", ".join(f"{k} {v}" for k, v in s["stall"].most_common(8)))
        print("   ops:", " ".join(f"{k}:{v / ctas:.0f}" for k, v in s["ops"].most_common(18)))
    allst = collections.Counter()
    for s in segs:
        allst.update(s["stall"])
    n = sum(allst.values())
    print("all stalls %:", ", ".join(f"{k} {100 * v / n:.1f}" for k, v in allst.most_common()))


if __name__ == "__main__":
    main()
