#!/bin/bash
# SASS evidence of the Blackwell-native paths (run here, no GPU): counts of tcgen05 / TMEM / TMA mnemonics per kernel family
# in the in-tree library.  usage: tools/sass_evidence.sh > profiles/sass_r02.txt
LIB=octave_b200/lib/liboctave_b200.so
echo "# cuobjdump -sass $LIB  (sm_100a cubins; $(date -u +%Y-%m-%dT%H:%MZ))"
echo "# UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG = TMA load, UTMASTG = TMA store, UTMAREDG = TMA reduce-add, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops"
cuobjdump -sass $LIB > /tmp/sass_all.txt
echo "whole library: $(grep -c 'UTCHMMA' /tmp/sass_all.txt) UTCHMMA, $(grep -c 'LDTM' /tmp/sass_all.txt) LDTM, $(grep -c 'UTMALDG' /tmp/sass_all.txt) UTMALDG, $(grep -c 'UTMASTG' /tmp/sass_all.txt) UTMASTG, $(grep -c 'UTMAREDG' /tmp/sass_all.txt) UTMAREDG, $(grep -c 'UTCBAR' /tmp/sass_all.txt) UTCBAR, $(grep -c 'HMMA\.' /tmp/sass_all.txt | tr -d '\n') legacy HMMA"
python3 - <<'PY'
import re, subprocess, collections
txt = open('/tmp/sass_all.txt').read()
parts = re.split(r'\n\s*Function : ', txt)
rows = []
for p in parts[1:]:
    name = p.split('\n', 1)[0].strip()
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r'\(anonymous namespace\)::', '', dem); dem = re.sub(r'\(.*$', '', dem)
    c = {k: len(re.findall(k, p)) for k in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTCBAR', 'UBLKCP')}
    if c['UTCHMMA'] or c['UTMALDG'] or c['UTMASTG']:
        rows.append((dem, c))
for dem, c in sorted(rows):
    print(f"{dem[:78]:78s} UTCHMMA {c['UTCHMMA']:3d}  LDTM {c['LDTM']:3d}  UTMALDG {c['UTMALDG']:3d}  UTMASTG {c['UTMASTG']:3d}  UTMAREDG {c['UTMAREDG']:3d}  UTCBAR {c['UTCBAR']:3d}")
PY
