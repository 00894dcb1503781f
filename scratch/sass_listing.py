"""profiles/r2z_sass_mnemonics.txt: SASS mnemonic counts per kernel of the built library (cuobjdump -sass)."""
import re, subprocess, sys
lib = "universal-beta-splatting_b200/ubs_b200/lib/libubs_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
parts = re.split(r'\n\s*Function : ', txt)
keys = ["UBLKCP", "SYNCS", "MATCH", "REDG", "REDS|RED\\.", "ATOMS", "ATOMG", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "FFMA",
        "HMMA|UTCMMA|UTCHMMA|QGMMA|IMMA"]
out = ["# SASS mnemonic counts per kernel of libubs_b200.so (sm_100a): `cuobjdump -sass <lib> | grep -c <mnemonic>` per function.",
       "# UBLKCP = TMA bulk copies (cp.async.bulk), SYNCS = mbarrier operations, MATCH = match.any, REDG / RED = reductions without",
       "# return (global atomics), ATOMS / ATOMG = returning shared / global atomics, MUFU = special-function unit.  The last column",
       "# counts tensor-core mnemonics: zero everywhere by design (no stage of the path is a dense contraction).",
       "# regenerate: python scratch/sass_listing.py > profiles/r2z_sass_mnemonics.txt", "",
       "%-58s %6s " % ("kernel", "instr") + " ".join("%6s" % k.split("|")[0][:6] for k in keys[:-1]) + " tensor"]
for p in parts[1:]:
    name = p.split('\n', 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r'ubs::\(anonymous namespace\)::', '', dem).split('(')[0].replace("void ", "")
    n_inst = len(re.findall(r'/\*[0-9a-f]{4,6}\*/\s+[A-Z@]', p))
    counts = [len(re.findall(r'\b(?:%s)' % k, p)) for k in keys]
    out.append("%-58s %6d " % (dem[:58], n_inst) + " ".join("%6d" % x for x in counts))
print("\n".join(out))
