#!/bin/bash
# Instruction histogram of libuic_b200.so (sm_100a SASS): the Blackwell-native evidence -- tcgen05 MMA (UTCHMMA), TMEM loads
# (LDTM), TMA tile / bulk copies (UTMALDG, UBLKCP), mbarrier ops (SYNCS), warp MMA of the attention context (HMMA).
# Usage: scripts/sass_histogram.sh > profiles/r2_sass_histogram.txt   (no GPU needed)
cd "$(dirname "$0")/.."
lib=unpaired_image_captioning_b200/libuic_b200.so
echo "# $(date -u +%Y-%m-%dT%H:%MZ)  $(nvcc --version | tail -1)"
echo "# cuobjdump -sass $lib | mnemonic counts (whole library, then per kernel for the tensor / TMA instructions)"
cuobjdump -sass $lib > /tmp/uic_sass.txt
grep -E '^\s+/\*[0-9a-f]{4,}\*/' /tmp/uic_sass.txt | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -60
echo
echo "# per kernel: UTCHMMA / LDTM / UTMALDG / UBLKCP / SYNCS / HMMA"
awk '/Function :/{fn=$3} /UTCHMMA/{a[fn]++} /LDTM/{b[fn]++} /UTMALDG/{c[fn]++} /UBLKCP/{d[fn]++} /SYNCS/{e[fn]++} / HMMA/{h[fn]++} END{for (f in e) printf "%6d %6d %6d %6d %6d %6d  %s\n", a[f], b[f], c[f], d[f], e[f], h[f], f}' /tmp/uic_sass.txt | sort -k7 | c++filt | cut -c1-200
