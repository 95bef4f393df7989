"""Per-kernel SASS opcode evidence (read here, no GPU): counts of the Blackwell-native mnemonics in every kernel of libvmae_sm100a.so.
usage: python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt
  UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA load/store/reduce, HMMA = legacy mma.sync (must be 0)."""
import re
import subprocess
from collections import OrderedDict
from pathlib import Path

lib = Path(__file__).resolve().parents[1] / "ssl_vit_video_analytics_b200" / "lib" / "libvmae_sm100a.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "MUFU.EX2", "REDG", "SYNCS"]
kernels = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = {o: 0 for o in ops}
        kernels[cur]["instructions"] = 0
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    kernels[cur]["instructions"] += 1
    op = m.group(1)
    for o in ops:
        if op.startswith(o) or (o == "HMMA" and op.startswith("HMMA")):
            kernels[cur][o] += 1


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*$", "", n).replace("void ", "").replace("vmae::", "") for n in out]


names = demangle(list(kernels))
print("# cuobjdump -sass libvmae_sm100a.so: instruction counts per kernel (sm_100a). tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG/UTMASTG/UTMAREDG")
print("# " + " | ".join(["kernel", "instructions"] + ops))
tot = {o: 0 for o in ops}
for (k, c), n in zip(kernels.items(), names):
    for o in ops:
        tot[o] += c[o]
    print(" | ".join([n[:110], str(c["instructions"])] + [str(c[o]) for o in ops]))
print("# total | - | " + " | ".join(str(tot[o]) for o in ops))
assert tot["HMMA"] == 0, "legacy mma.sync found"
