"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass): the instructions that prove what the
kernels run on -- DMMA (FP64 tensor cores), UTMALDG (TMA tensor loads), UBLKCP (bulk copies), SYNCS (mbarriers), DFMA,
LDS/STS, ATOM.  Writes profiles/sass_summary_r02.json.   python tools/sass_summary.py"""
import collections, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "libdmet_preview_b200", "libldm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, counts = None, {}
keys = ("DMMA", "UTMALDG", "UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "LDS", "STS", "LDG", "STG", "ATOM", "ATOMG", "RED", "BAR", "SHFL",
        "UTCHMMA", "UTCQMMA", "LDTM", "HMMA")
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["_total"] += 1
        if op in keys:
            counts[kern][op] += 1
res = {k: dict(v) for k, v in sorted(counts.items())}
total = collections.Counter()
for v in res.values():
    total.update(v)
json.dump({"library": "libdmet_preview_b200/libldm_b200.so", "arch": "sm_100a", "totals": dict(total), "kernels": res},
          open(os.path.join(ROOT, "profiles", "sass_summary_r02.json"), "w"), indent=1)
print(json.dumps(dict(total)))
for k, v in res.items():
    if v.get("DMMA") or v.get("UTMALDG") or v.get("UBLKCP"):
        print("%-70s %s" % (k[:70], {x: v[x] for x in ("DMMA", "UTMALDG", "UBLKCP", "SYNCS") if v.get(x)}))
