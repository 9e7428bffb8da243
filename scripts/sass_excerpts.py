"""profiles/sass_<round>_excerpts.txt: per kernel of libdvfe.so, the SASS opcode histogram and every asynchronous-copy / TMA / mbarrier /
warp-reduction instruction (the mnemonics B200_PROFILING.md names as proof of TMA, cp.async and redux use).
usage: python scripts/sass_excerpts.py > profiles/sass_r2_excerpts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "dynamic_vins_b200", "libdvfe.so")], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", ln)
    if m and cur:
        funcs[cur].append((m.group(1), m.group(2).strip()))
SPECIAL = ("UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "REDUX", "UBLKCP", "LDTM", "STTM", "UTCHMMA", "HMMA")
print("cuobjdump -sass dynamic_vins_b200/libdvfe.so (sm_100a): opcode counts of every kernel and each asynchronous-copy / TMA / mbarrier /\n"
      "warp-reduction instruction in it.  UTMALDG = cp.async.bulk.tensor (TMA tile load), SYNCS = mbarrier, LDGSTS = cp.async,\n"
      "REDUX = redux.sync, IDP = dp4a / dp2a.  No tensor-core instruction (HMMA / UTCHMMA / LDTM) exists: nothing on this path is a\ndense contraction.\n")
tot = collections.Counter()
for name, ins in funcs.items():
    ops = collections.Counter()
    for _, t in ins:
        tok = t.split()
        op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
        ops[op.split(".")[0]] += 1
    tot.update(ops)
    short = re.sub(r"^_ZN?\d*_?GLOBAL__N__\w+?_cu_\w{8}\d+", "", name)
    print(f"{name}: {len(ins)} SASS instructions; " + ", ".join(f"{o} {c}" for o, c in ops.most_common(12)))
    for key in SPECIAL + ("IDP", "SHFL", "DADD"):
        if ops.get(key):
            print(f"    {key}: {ops[key]}")
    for addr, t in ins:
        if any(t.split()[-0].startswith(k) or (t.startswith("@") and t.split()[1].startswith(k)) for k in ("UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "UBLKCP")):
            print(f"      {addr} {t} ;")
    print()
print("whole library: " + ", ".join(f"{k} {tot.get(k, 0)}" for k in ("UTMALDG", "SYNCS", "LDGSTS", "REDUX", "IDP", "SHFL", "HMMA", "UTCHMMA", "LDTM")))
