#!/usr/bin/env python
"""SASS mnemonic counts per kernel instance from the objects liborion_b200.so is linked from (cuobjdump -sass; runs
without a GPU).  Usage: python scripts/sass_table.py > profiles/rNN_sass_mnemonics.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "orion-sdr_b200", "build")
UNITS = ["chain_inst_hot", "chain_inst_direct", "chain_inst_staged_u1", "chain_inst_staged_u2", "chain_inst_ws", "chain_launch",
         "bank_kernels", "agc_kernels", "aux_kernels"]
COLS = [("UTMALDG", r"UTMALDG"), ("UBLKCP", r"UBLKCP"), ("SYNCS", r"SYNCS"), ("FFMA2", r"FFMA2"), ("FMUL2", r"FMUL2"),
        ("FFMA", r"FFMA(?!2)"), ("SHFL", r"SHFL"), ("LDL+STL", r"\b(LDL|STL)\b"), ("USETMAXREG", r"USETMAXREG"),
        ("tensor", r"UTC[A-Z]*MMA|\bLDTM\b|\bSTTM\b|\bHMMA\b|\bIMMA\b")]
print("SASS mnemonic counts per kernel instance (cuobjdump -sass of the objects liborion_b200.so is linked from; sm_100a).")
print("UTMALDG = cp.async.bulk.tensor (TMA load), UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, FFMA2 / FMUL2 = packed f32x2,")
print("USETMAXREG = setmaxnreg; tensor-core mnemonics (UTCMMA / UTCHMMA / LDTM / STTM) would show up in the last column.\n")
print(f"{'kernel':78s} {'instr':>7s} " + " ".join(f"{c:>7s}" for c, _ in COLS))
for u in UNITS:
    obj = os.path.join(BUILD, u + ".o")
    if not os.path.exists(obj):
        continue
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, body = None, []
    def flush():
        if name is None: return
        ins = [l for l in body if re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", l)]
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        dem = re.sub(r"^void ", "", dem).replace("orion::", "").replace("(int)", ""); dem = re.sub(r"\((?:ChainArgs|BankFirArgs|const|AgcArgs|[A-Za-z]).*$", "", dem)
        print(f"{(u + ': ' + dem)[:78]:78s} {len(ins):7d} " + " ".join(f"{sum(1 for l in ins if re.search(rx, l)):7d}" for _, rx in COLS))
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            flush(); name, body = m.group(1), []
        else:
            body.append(line)
    flush()
