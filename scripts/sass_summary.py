"""SASS / resource summary of the hot kernels of the built library -> profiles/r2_sass_summary.md
(cuobjdump -sass + --dump-resource-usage; run on the CPU box after __graft_entry__.build())."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "goldfish_b200", "libgoldfish_b200.so")
HOT = ["k_spmv_node", "k_spmv_simple", "k_sw_solve1", "k_sw_coarse_cluster", "k_shell_k2", "k_shell_p2", "k_sw_update", "k_sw_potrf",
       "k_dense_rows", "k_pcg_update", "k_penalty_points", "k_peak_dmma"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1); continue
    if cur and "REG:" in line:
        usage[cur] = line.strip(); cur = None
blocks = re.split(r"\n\s*Function : ", sass)
out = ["# SASS summary of the hot kernels (sm_100a, cuobjdump -sass of goldfish_b200/libgoldfish_b200.so)", "",
       "Instruction counts are static (per kernel body), mnemonics grouped; `LDG.E.128` = 128-bit global loads, `DFMA`/`DMUL`/`DADD` = FP64 pipe,",
       "`DMMA` = FP64 tensor-core MMA (only in the peak micro-benchmark: the element contraction stays on DFMA, see DESIGN.md), `UCGABAR`/`CCTL`/`MEMBAR` = cluster and memory barriers.", ""]
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    short = next((h for h in HOT if h in name), None)
    if short is None:
        continue
    ops = collections.Counter()
    for line in b.splitlines():
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1)] += 1
    tot = sum(ops.values())
    grp = collections.Counter()
    for k, v in ops.items():
        base = k.split(".")[0]
        if k.startswith("LDG.E.128") or k.startswith("LDG.E.LTC128B.128") or ".128" in k and base == "LDG":
            grp["LDG.*.128"] += v
        elif base in ("DFMA", "DMUL", "DADD", "DMMA", "DSETP", "MUFU", "LDG", "STG", "LDS", "STS", "BAR", "SHFL", "FFMA", "HMMA", "UCGABAR_ARV", "UCGABAR_WAIT", "LDGSTS", "LDGDEPBAR", "CCTL", "MEMBAR", "ATOM", "ATOMG", "RED", "PRMT", "IMAD", "LDSM", "UTMALDG", "SYNCS"):
            grp[base] += v
    u = next((v for k, v in usage.items() if short in k and name[:40] in k), None) or next((v for k, v in usage.items() if name in k), "")
    out.append("## `%s`" % name[:110])
    out.append("")
    out.append("%d instructions; %s" % (tot, u))
    out.append("")
    out.append(", ".join("%s %d" % kv for kv in sorted(grp.items(), key=lambda kv: -kv[1])))
    out.append("")
open(os.path.join(ROOT, "profiles", "r2_sass_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
