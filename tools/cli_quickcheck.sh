#!/bin/bash
# Fast GPU check of the CLI without Python: golden -hist (with and without -seqmers) and -completeness
# without -seqmers against the reference's committed outputs.  Prints PASS/FAIL lines.
cd "$(dirname "$0")/.." || exit 1
M=merfin_b200/bin/merfin
t=$(mktemp -d)
rc=0
for spec in g1_k21:8 g2_k6:3 g3_k31:6; do
  n=${spec%%:*}; p=${spec##*:}; d=tests/golden/$n
  $M -hist -sequence $d/asm.fa -seqmers $d/asm.meryl -readmers $d/reads.meryl -peak $p -output $t/a.hist 2> $t/a.err \
    && cmp -s $t/a.hist $d/ref.hist && echo "PASS $n hist" || { echo "FAIL $n hist"; rc=1; }
  $M -hist -sequence $d/asm.fa -readmers $d/reads.meryl -peak $p -output $t/b.hist 2> $t/b.err \
    && cmp -s $t/b.hist $d/ref.hist && echo "PASS $n hist (no -seqmers)" || { echo "FAIL $n hist (no -seqmers)"; rc=1; }
  $M -completeness -sequence $d/asm.fa -readmers $d/reads.meryl -peak $p 2> $t/c.err
  ok=1
  while IFS= read -r l; do grep -qxF -- "$l" $t/c.err || ok=0; done < <(grep -v '^thread' $d/ref.completeness | grep -v '^$')
  [ $ok = 1 ] && echo "PASS $n completeness (no -seqmers)" || { echo "FAIL $n completeness (no -seqmers)"; tail -5 $t/c.err; rc=1; }
done
rm -rf $t
exit $rc
