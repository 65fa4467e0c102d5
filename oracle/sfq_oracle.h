/* TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of slimfastq's entropy-coding hot path (one FASTQ buffer in, the
 * reference's named byte streams + info keys out, and back).  Only tests/, bench.py's
 * cpu_baseline leg / --impl reference arm and __graft_entry__.smoke() may load this; the product
 * library (slimfastq_b200/csrc) never links, includes or calls it.
 *
 * Parity status: PINNED.  The reference ships no byte-level golden vectors (its tests are
 * round-trips only), so the pin is the unmodified reference binary built by
 * oracle/Makefile into oracle/_ref/slimfastq: tests/test_oracle.py compares every stream of
 * this restatement with the streams extracted from the reference's own .sfq output on all 16
 * reference samples x levels 1-4 and on the synthetic edge-case corpus, and
 * tests/golden/golden.json holds the reference's stream md5s for inputs that
 * tests/golden/make_golden.py regenerates deterministically.
 */
#ifndef SFQ_ORACLE_H
#define SFQ_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Stream ids, in the order the reference may create them (usrs.cpp:47-54, 396-398;
 * gens.cpp:68-69; recs.cpp:46). */
enum {
    SFQ_OR_REC = 0, SFQ_OR_GEN, SFQ_OR_QLT, SFQ_OR_GEN_NS, SFQ_OR_GEN_NN, SFQ_OR_REC_X,
    SFQ_OR_USR_X, SFQ_OR_USR_XQ, SFQ_OR_USR_PFG, SFQ_OR_USR_PFQ,
    SFQ_OR_USR_LREC, SFQ_OR_USR_LGEN, SFQ_OR_USR_LQLT,      /* oversized records, usrs.cpp:269-301 */
    SFQ_OR_NSTREAMS
};

typedef struct sfq_or_chunk {
    /* info keys of the reference's stream #0 that carry meaning (config.cpp:334-336;
     * usrs.cpp:262-266, 405; recs.cpp:68-75; gens.cpp:100-105; qlts.cpp:57-60) */
    int32_t  level;          /* config.level, 1..4                                  */
    int32_t  llen;           /* llen: first record's base-line length (-1 if SOLiD)  */
    int32_t  solid;          /* usr.solid                                            */
    int32_t  two_id;         /* usr.2id                                              */
    int32_t  n_byte;         /* gen.N_byte, 0 when the key is absent ('N' implied)   */
    uint32_t extra_hi_qlt;   /* qlt.extra.hi (count of qualities >= 63)              */
    uint64_t num_records;    /* num_records                                          */
    int32_t  version;        /* info key `version` (0 = absent); < 5 selects the pre-v5 header stream, recs.cpp:397-398 */
    char     rec_first[0x200];
    uint32_t rec_first_len;
    /* streams; size 0 and data NULL = the reference would not have created the stream */
    uint8_t *data[SFQ_OR_NSTREAMS];
    size_t   size[SFQ_OR_NSTREAMS];
} sfq_or_chunk;

const char *sfq_oracle_stream_name(int id);

/* Encode one FASTQ buffer exactly as `slimfastq -u buf -f out -l level -q` would
 * (UsrSave::encode, usrs.cpp:392-407).  Returns 0, or non-zero with a croak-style message in
 * err[256].  Oversized records (usrs.hpp:34-36: id of 8 KiB or line of 64 KiB) go to usr.lrec / usr.lgen / usr.lqlt. */
int sfq_oracle_encode(const uint8_t *fastq, size_t n, int level, sfq_or_chunk *out, char *err);
/* Same, writing the header stream the way a pre-v5 slimfastq did (what RecLoad::load_pre5, recs.cpp:463-510, reads:
 * decimal fields as DGT/DLT gaps against the previous header's TEXT, everything else as strings).  The reference
 * ships no pre-v5 encoder; this one exists so that the legacy decode path can be tested, and it is pinned the other
 * way round: the reference binary must decode what it writes (tests/test_legacy.py). */
int sfq_oracle_encode_pre5(const uint8_t *fastq, size_t n, int level, sfq_or_chunk *out, char *err);

/* Decode (UsrLoad::decode, usrs.cpp:539-574).  *out is malloc'ed. */
int sfq_oracle_decode(const sfq_or_chunk *in, uint8_t **out, size_t *out_n, char *err);

void sfq_oracle_free_chunk(sfq_or_chunk *c);
void sfq_oracle_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
