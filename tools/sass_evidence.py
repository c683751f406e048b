"""Per-kernel counts of the tcgen05 / TMEM / TMA / cluster SASS instructions in the built library.
Usage: python tools/sass_evidence.py > profiles/rNN_sass_evidence.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "instaorder_b200", "libinstaorder_b200.so")
PAT = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCBAR(?:\.2CTA\.MULTICAST)?|LDTM|UTMALDG\.\dD(?:\.2CTA)?|UTMASTG\.\dD|UCGABAR_ARV|"
                 r"STG\.E\.ENL2\.256|SYNCS)\b")
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "UTCBAR.2CTA.MULTICAST", "LDTM", "UTMALDG", "UTMALDG.2CTA", "UTMASTG",
        "UCGABAR_ARV", "STG.E.ENL2.256", "SYNCS"]


def demangle(n):
    try:
        s = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
        return s.split("(")[0].replace("io::", "").replace("void ", "")
    except Exception:
        return n


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    kern, cnt = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = m.group(1)
            cnt[kern] = collections.Counter()
            continue
        if kern:
            for t in PAT.findall(line):
                cnt[kern][re.sub(r"\.\dD", "", t)] += 1
    print("# SASS evidence (`cuobjdump -sass instaorder_b200/libinstaorder_b200.so`, sm_100a): tcgen05 / TMEM / TMA / cluster "
          "instructions per kernel\n")
    print("`UTCHMMA` = tcgen05.mma (`.2CTA` = `cta_group::2`, issued by the leader CTA of a pair), `UTCBAR` = tcgen05.commit -> "
          "mbarrier (`.2CTA.MULTICAST` = arrive in both CTAs of the pair), `LDTM` = tcgen05.ld (TMEM -> registers), `UTMALDG` / "
          "`UTMASTG` = TMA tensor load / store (`.2CTA` = bytes counted on the leader CTA's barrier), `UCGABAR_ARV` = "
          "barrier.cluster.arrive, `STG.E.ENL2.256` = 256-bit global store, `SYNCS` = mbarrier operations.\n")
    print("| kernel | " + " | ".join(COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for k, c in cnt.items():
        if c["UTCHMMA"] + c["UTCHMMA.2CTA"] + c["UTMALDG"] + c["UTMALDG.2CTA"] + c["STG.E.ENL2.256"] == 0:
            continue
        print("| `%s` | " % demangle(k) + " | ".join(str(c[x]) for x in COLS) + " |")


if __name__ == "__main__":
    main()
