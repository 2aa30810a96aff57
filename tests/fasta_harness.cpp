// Test harness: prints what the product's FASTA/FASTQ loader (merfin_b200/csrc/host/fasta.cpp)
// makes of a file, one "ident<TAB>length<TAB>bases" line per record.  Test infrastructure only.
#include <cstdio>
#include <string>

#include "../merfin_b200/csrc/host/fasta.h"

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  SeqSet S; std::string err;
  if (!load_sequences(argv[1], S, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
  for (size_t i = 0; i < S.nseq(); i++) {
    printf("%s\t%lu\t", S.ident[i].c_str(), (unsigned long)S.length(i));
    fwrite(S.seq(i), 1, S.length(i), stdout);
    putchar('\n');
  }
  return 0;
}
