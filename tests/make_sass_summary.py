"""Static evidence from the shipped library, made WITHOUT a GPU (cuobjdump on lib/libqat_b200.so):
per kernel family the registers / stack / static shared memory (`cuobjdump -res-usage`) and the SASS
mnemonics that prove which hardware path a kernel uses (`cuobjdump -sass`): UTCIMMA / UTCHMMA
(tcgen05.mma kind::i8 / kind::f16), UTCBAR (tcgen05.commit), LDTM / STTM (tcgen05.ld / st),
UTMALDG / UTMASTG (TMA tensor loads / stores), SYNCS (mbarrier), MUFU.EX2, HMMA (none expected:
no mma.sync anywhere), and local-memory LDL / STL (spills).

    python tests/make_sass_summary.py > profiles/r02_sass_summary.json
"""
from __future__ import annotations

import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "llm-qat_b200", "lib", "libqat_b200.so")
WATCH = ["UTCIMMA", "UTCHMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS",
         "MUFU.EX2", "MUFU.RCP", "HMMA", "IMMA", "LDL", "STL", "ACQBULK", "LDGDEPBAR", "BAR.SYNC", "ATOM", "RED"]


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True)
    return r.stdout.splitlines()


def short(name: str) -> str:
    """qat::(anonymous namespace)::kernel<args>(params) -> kernel<args>"""
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    depth = 0
    for i, ch in enumerate(name):          # cut the parameter list (first '(' outside <...>)
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            name = name[:i]
            break
    return name.replace("qat::", "")


def main():
    if not os.path.exists(LIB):
        sys.exit(f"{LIB} missing: run python llm-qat_b200/build.py first")
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in ln:
            f = dict(re.findall(r"(\w+(?:\[\d\])?):(\d+)", ln))
            usage[cur] = {"reg": int(f["REG"]), "stack": int(f["STACK"]), "static_smem": int(f["SHARED"]),
                          "local": int(f["LOCAL"])}
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.defaultdict(collections.Counter)
    ninstr = collections.Counter()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        ninstr[cur] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w in ("LDL", "STL") and op.startswith(w)):
                counts[cur][w] += 1
    names = sorted(usage)
    pretty = dict(zip(names, (short(n) for n in demangle(names))))
    # family = kernel name without template arguments
    fam = collections.defaultdict(list)
    for n in names:
        fam[pretty[n].split("<")[0]].append(n)
    out = {"what": __doc__.split("\n\n")[0].replace("\n", " "),
           "library": os.path.relpath(LIB, ROOT), "kernels_total": len(names), "families": {}}
    arch = set(re.findall(r"arch = (sm_\w+)", sass))
    out["arch"] = sorted(arch)
    for f, members in sorted(fam.items()):
        regs = [usage[n]["reg"] for n in members]
        tot = collections.Counter()
        for n in members:
            tot.update(counts[n])
        worst = max(members, key=lambda n: (usage[n]["stack"], usage[n]["reg"]))
        entry = {"instances": len(members), "reg_min": min(regs), "reg_max": max(regs),
                 "stack_max": max(usage[n]["stack"] for n in members),
                 "sass_instructions_max": max(ninstr[n] for n in members),
                 "mnemonics_all_instances": {k: v for k, v in sorted(tot.items()) if v}}
        if usage[worst]["stack"]:
            entry["largest_stack_instance"] = {"name": pretty[worst], **usage[worst],
                                               "LDL": counts[worst]["LDL"], "STL": counts[worst]["STL"]}
        out["families"][f] = entry
    whole = collections.Counter()
    for c in counts.values():
        whole.update(c)
    out["mnemonics_whole_library"] = {k: whole[k] for k in WATCH}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
