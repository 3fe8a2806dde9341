"""Opcode evidence per kernel of libb200vq.so: which kernels carry tcgen05 / TMEM / TMA / multimem instructions.
    python tools/sass_summary.py [path/to/libb200vq.so] > profiles/r2_sass_summary.txt
SASS mnemonics (B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA tensor loads -> UTMALDG,
bulk copies -> UBLKCP, tcgen05.commit -> UTCBAR, red.global -> RED / REDG, mbarrier -> SYNCS."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "acoustic_locating_vq-vae_b200", "csrc", "libb200vq.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "RED", "REDG", "ATOMG", "ATOMS",
         "HMMA", "HGMMA", "MULTIMEM", "LDGSTS", "ELECT", "UCGABAR_ARV", "UCGABAR_WAIT", "ACQBULK", "ST", "LD"]
fn = None
counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and fn:
        op, mods = m.group(1), m.group(2)
        counts[fn]["_total"] += 1
        if op in WATCH:
            counts[fn][op] += 1
        if op == "UTCHMMA" and ".2CTA" in mods:
            counts[fn]["UTCHMMA.2CTA"] += 1
        if "MULTIMEM" in mods or "MMIO" in mods:
            counts[fn]["mods:" + mods] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# opcode evidence per kernel of {os.path.relpath(so, ROOT)} (cuobjdump -sass; tools/sass_summary.py)")
print("# columns: SASS instructions | tcgen05.mma (of which cta_group::2) | tcgen05.ld | TMA tensor loads | bulk copies | tcgen05.commit | mbarrier ops | red / atom")
for (fnm, c), name in zip(counts.items(), demangle):
    short = re.sub(r"\(.*", "", name).replace("b200vq::", "").replace("void ", "")
    extra = " ".join(f"{k}={v}" for k, v in c.items() if k.startswith("mods:"))
    print(f"{short:60s} {c['_total']:6d} | UTCHMMA {c['UTCHMMA']:3d} (2CTA {c['UTCHMMA.2CTA']:3d}) | LDTM {c['LDTM']:3d} | UTMALDG {c['UTMALDG']:3d} | UBLKCP {c['UBLKCP']:3d} | "
          f"UTCBAR {c['UTCBAR']:3d} | SYNCS {c['SYNCS']:3d} | RED {c['RED'] + c['REDG']:3d} ATOMG {c['ATOMG']:3d} ATOMS {c['ATOMS']:3d} HMMA {c['HMMA']} {extra}")
