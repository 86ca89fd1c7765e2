#!/bin/bash
# Count the Blackwell-specific SASS mnemonics in the objects of libairv2x_b200.so (run after build(); no GPU needed):
#   UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor loads, LDTM / STTM = tcgen05.ld / .st (tensor memory), UTCBAR = tcgen05.commit,
#   UTCATOMSWS = tcgen05.alloc / dealloc, SYNCS = mbarrier operations.   scripts/sass_mnemonics.sh > profiles/rN_sass_mnemonics.md
cd "$(dirname "$0")/../airv2x-perception_b200/build" || exit 1
echo "| object | UTCHMMA | UTMALDG | LDTM | STTM | UTCBAR | UTCATOMSWS | SYNCS (mbarrier) |"
echo "|---|---|---|---|---|---|---|---|"
for o in *.o; do
  s=$(cuobjdump -sass "$o" 2>/dev/null)
  c() { echo "$s" | grep -cE "\b$1"; }
  echo "| $o | $(c UTCHMMA) | $(c UTMALDG) | $(c LDTM) | $(c STTM) | $(c UTCBAR) | $(c UTCATOMSWS) | $(c SYNCS) |"
done
