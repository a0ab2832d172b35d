"""SASS opcode summary of the built library → profiles/rNN_sass_summary.txt (run on the build box, no GPU needed):
    python tools/sass_summary.py profiles/r02_sass_summary.txt
Proves which tensor-core / copy paths the kernels use (UTCHMMA = tcgen05.mma, UTMALDG = TMA, LDTM/STTM = tcgen05.ld/st)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "art_sbir_b200" / "lib" / "libsbir_b200.so"
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "ELECT", "HMMA", "HGMMA", "UBLKCP", "UBLKPF",
       "LDGSTS", "ATOMG", "BAR.SYNC", "MEMBAR", "DFMA", "FMNMX3"]


def main(out_path):
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    elf = subprocess.run(["cuobjdump", "-lelf", str(LIB)], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    tot, kinds = collections.Counter(), collections.Counter()
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        c = collections.Counter()
        for line in f.split("\n"):
            m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + ".") or (o == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                    c[o] += 1
        tot.update(c)
        if c.get("UTCHMMA") or c.get("UTMALDG") or c.get("LDTM"):
            kinds["dist_topk_kernel" if "dist_topk_kernel" in name else "batch_hard_fused_kernel" if "batch_hard" in name else name[:60]] += 1
    lines = ["SASS opcode summary of art_sbir_b200/lib/libsbir_b200.so, `cuobjdump -sass`, counted per opcode over all kernels",
             "ELF images: " + ", ".join(sorted(set(re.findall(r"sm_\w+", elf)))), f"kernel functions: {len(funcs) - 1}", ""]
    lines += [f"{o:14s} {tot.get(o, 0)}" for o in OPS]
    lines += ["", "UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA tensor load (cp.async.bulk.tensor), LDTM / STTM = tcgen05.ld / st,",
              "UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, ELECT = elect.sync; HMMA / HGMMA (mma.sync / Hopper wgmma) must be 0.", "",
              "kernels that contain tcgen05 / TMA instructions (distinct template instantiations):"]
    lines += [f"  {k}: {v}" for k, v in kinds.items()]
    Path(out_path).write_text("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "profiles" / "sass_summary.txt"))
