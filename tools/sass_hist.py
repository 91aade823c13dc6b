import subprocess, re, collections
out = subprocess.run(["cuobjdump", "-sass", "desmo_b200/libdesmo_b200.so"], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "UTCATOMSWS", "FFMA", "F2FP", "LDS", "STS", "LDG", "STG"]
kern, cnt, order = None, {}, []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); cnt[kern] = collections.Counter(); order.append(kern); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1); cnt[kern]["total"] += 1
        for o in ops:
            if op == o or op.startswith(o + "."):
                cnt[kern][o] += 1
def dem(n):
    s = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()
    s = s.replace("(bool)0", "false").replace("(bool)1", "true")
    s = re.sub(r"\(int\)(\d+)", r"\1", s)
    return re.sub(r"\(.*", "", s)[:64]
print("SASS opcode counts per kernel (cuobjdump -sass desmo_b200/libdesmo_b200.so, sm_100a, final build of round 2) -- the Blackwell-native instructions:")
print("UTCHMMA = tcgen05.mma kind::f16, UTMALDG = TMA tensor load, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops")
print(f"{'kernel':64s} " + " ".join(f"{o:>8s}" for o in ops) + f" {'total':>8s}")
for k in sorted(order, key=lambda k: (-cnt[k]["UTCHMMA"], -cnt[k]["total"])):
    print(f"{dem(k):64s} " + " ".join(f"{cnt[k][o]:8d}" for o in ops) + f" {cnt[k]['total']:8d}")
