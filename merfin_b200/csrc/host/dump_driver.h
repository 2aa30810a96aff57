// -dump driver of the merfin-compatible front-end (merfin-dump.C:20-104): the records of a sequence set
// are cut into batches; the (v_read, v_asm, valid) triples of one batch are computed by a callback
// (the CLI passes mfn_dump) while the previous batch is formatted and written by mfn_dump_write.
// Host memory is two batches of 9 bytes per base, reused, instead of 9 bytes per base of the whole
// assembly (27 GB at 3 Gbp, whose first touch alone costs seconds).
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>

#include "../../../include/merfin_b200.h"
#include "fasta.h"

// bases = the batch's records concatenated, off = nrec + 1 offsets starting at 0; fills v_read / v_asm /
// valid for off[nrec] positions; returns MFN_OK or an error status.
typedef std::function<int(const char *bases, const uint64_t *off, uint32_t nrec, uint32_t *v_read, uint32_t *v_asm, uint8_t *valid)>
    DumpBatchFn;

// Writes the -dump lines of every record to fd (fd < 0: counts only, as -skipMissing does) and one
// "ident <tab> missing <tab> cumulative missing <tab> cumulative found" line per record to `log`
// (merfin-dump.C:100-101).  batch_bases: records are grouped until a batch holds at least this many
// bases (0 = default: MFN_DUMP_BATCH or 32 M).  Returns MFN_OK or the first error status (detail in *err).
int dump_records(const SeqSet &S, const mfn_kmodel *km, int fd, int threads, uint64_t batch_bases,
                 const DumpBatchFn &pairs, FILE *log, std::string *err);
