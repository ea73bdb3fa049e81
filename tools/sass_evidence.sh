#!/usr/bin/env bash
# Regenerates profiles/r02_sass_evidence.txt: Blackwell-specific SASS mnemonics of the in-tree objects (build first).
set -euo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
B=eeg-gan-timegan-cgan_b200/csrc/build
for f in proj_tcgen05 proj_bf16 wgrad_gru_tcgen05 wgrad_tcgen05 gru_fwd gru_bwd gru_jvp gru_cluster; do
  echo "== $f.o"
  cuobjdump -sass $B/$f.o | grep -oE "\b(UTCHMMA|LDTM|UTMALDG[.A-Z0-9]*|UTMASTG[.A-Z0-9]*|UBLKCP[.A-Z0-9]*|SYNCS[.A-Z0-9]*|FFMA2|STAS[.A-Z0-9]*|UCGABAR_[A-Z]*|LDGSTS[.A-Z0-9]*|MUFU\.[A-Z0-9]*|SHFL\.[A-Z]*|UTCBAR[.A-Z0-9]*)" | sort | uniq -c | sort -rn | head -12
done
