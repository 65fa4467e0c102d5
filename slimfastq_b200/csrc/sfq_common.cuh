// Shared definitions for the sm_100a kernels of the slimfastq hot path.
//
// Every per-chunk coder routine is written as an SFQ_HD function over plain pointers so that
// (a) the kernels are thin wrappers and (b) tests/emul can compile the very same routines with
// g++ and single-step them on the CPU-only dev container.  The emulation build is test tooling:
// the shipped library (sfq_abi.cu) only ever launches the __global__ wrappers and refuses to
// work without a CUDA device.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define SFQ_HD __host__ __device__ __forceinline__
#define SFQ_HDN static __host__ __device__
// rare paths are kept out of line so the hot loops stay small enough for the instruction caches
#define SFQ_COLD __host__ __device__ __noinline__
#else
#define SFQ_HD inline
#define SFQ_HDN static
#define SFQ_COLD
#endif

// Stream ids inside a chunk.  Names/ordering follow the reference's stream creation sites
// (usrs.cpp:47-54,396-398; gens.cpp:68-69; recs.cpp:46).
enum {
    SFQ_S_REC = 0, SFQ_S_GEN, SFQ_S_QLT, SFQ_S_GEN_NS, SFQ_S_GEN_NN, SFQ_S_REC_X,
    SFQ_S_USR_X, SFQ_S_USR_XQ, SFQ_S_USR_PFG, SFQ_S_USR_PFQ,
    SFQ_S_USR_LREC, SFQ_S_USR_LGEN, SFQ_S_USR_LQLT,        // oversized records, stored verbatim (usrs.cpp:269-301)
    SFQ_NSTREAMS
};

// Per-chunk status codes written by kernels; the host turns the first non-zero one into the
// reference's croak text (config.cpp:54-68) and a non-zero return.
enum {
    SFQ_OK = 0,
    SFQ_E_AT = 1,          // record does not start with '@'            usrs.cpp:158-163,200
    SFQ_E_PLUS = 2,        // third line does not start with '+'        usrs.cpp:232,346
    SFQ_E_TRUNC = 3,       // truncated record / missing final newline  usrs.cpp:165-168
    SFQ_E_OVERSIZE = 4,    // '+' line of 8 191+ chars ("wierd second id", usrs.cpp:349-352); oversized records themselves are coded
    SFQ_E_BASE = 5,        // unexpected genome char                    gens.cpp:125-126
    SFQ_E_NBYTE = 6,       // switched N byte                           gens.cpp:107-108
    SFQ_E_SEPS = 7,        // > 64 separators in a header               recs.cpp:153-154
    SFQ_E_CAP = 8,         // a stream outgrew its device arena (host retries with more room)
    SFQ_E_TABLE = 9,       // base-context hash table full (host retries with a larger table)
    SFQ_E_FIRSTHDR = 10,   // first header > 399 chars                  recs.cpp:31,69-70
    SFQ_E_CORRUPT = 11,    // decoder: impossible value in a stream
    SFQ_E_EMPTYSEQ = 12,   // first record of a chunk has an empty base line (usrs.cpp:216-231 cannot represent it)
    SFQ_E_CHUNKSIZE = 13,  // a chunk's bases, qualities or headers do not fit the container's 32-bit counts
};

// Sizes of the per-chunk model pools.
#define SFQ_L64_WORDS   64u            // one quality context  = 64 packed slots = 256 B
#define SFQ_PW_WORDS    340u           // one 256-symbol model: 256 slots + inverse map + group sums (SfqPower)
// 256-symbol model instances per chunk:
//   header fields: 66 x (type, str, num[14])            recs.hpp:42-48
//   10 exception streams x (num[14], str)               xfile.hpp:41-42
//   1 quality escape model                              qlts.hpp:44
#define SFQ_PW_REC_BASE   0u
#define SFQ_PW_PER_FIELD  16u
#define SFQ_PW_X_BASE     (66u * 16u)
#define SFQ_PW_PER_X      15u
#define SFQ_PW_QEX        (SFQ_PW_X_BASE + 10u * SFQ_PW_PER_X)
#define SFQ_PW_PER_CHUNK  (SFQ_PW_QEX + 1u)      // 1207 models = 1.6 MiB per resident chunk
// exception-stream slot -> stream id
#define SFQ_X_NS 0
#define SFQ_X_NN 1
#define SFQ_X_REC 2
#define SFQ_X_LLEN 3
#define SFQ_X_QLEN 4
#define SFQ_X_SGEN 5
#define SFQ_X_SQLT 6
#define SFQ_X_LREC 7
#define SFQ_X_LGEN 8
#define SFQ_X_LQLT 9

#define SFQ_MAX_ID_LLEN 0x2000
#define SFQ_MAX_GN_LLEN 0x10000

// What the host learns about a chunk from the planning kernel, and what every coder needs.
struct SfqChunkMeta {
    uint64_t text_off;      // byte offset of the chunk's first record in the FASTQ buffer
    uint64_t text_len;      // bytes of FASTQ text in the chunk
    uint64_t out_len;       // bytes the decoder will print for it (== text_len unless the input
                            // has one of the reference's lossy cases, e.g. a 2nd id that differs)
    uint64_t line0;         // index of the chunk's first line in the line-start table
    uint32_t nrec;          // records in the chunk (num_records, usrs.cpp:405)
    uint32_t nbases;        // sum of coded base-line lengths
    uint32_t nquals;        // sum of coded quality-line lengths
    uint32_t hdr_bytes;     // sum of header lengths (without '@' and '\n')
    int32_t  llen;          // `llen` info key: first record's base-line length (usrs.cpp:265)
    uint8_t  solid;         // usr.solid  (usrs.cpp:242-263)
    uint8_t  two_id;        // usr.2id    (usrs.cpp:235-239,266)
    uint8_t  n_byte;        // gen.N_byte (gens.cpp:100-105); 0 = key absent
    uint8_t  pad;
    uint32_t extra_hi;      // qlt.extra.hi (qlts.cpp:57-60)
    uint32_t q_used;        // distinct quality contexts the chunk touched (sizes the decoder's table)
    uint32_t g_used;        // distinct base contexts the chunk touched
    uint32_t status;        // SFQ_OK or first error
    uint32_t status_arg;    // record number / offending byte for the message
    // oversized records of the chunk (usrs.cpp:269-301): how many, and the bytes of their lines - base lines and quality
    // lines without the newline, id line (without '@') + '\n' + '+' line as one block.  nbases / nquals / hdr_bytes count
    // the CODED records only; the decoder's planes hold both.
    uint32_t nbig, big_bases, big_quals, big_hdr;
    uint32_t first_coded;   // index of the first record that goes through the models (its id line is `rec.first`); nrec if none
    uint32_t pad2;
};

// Bytes reserved for a chunk's decoded headers (each followed by '\n').  The slack covers the few
// cases where the reference prints a header field longer than it was (sign of a "%lld" value).
#define SFQ_HDR_PLANE(m) ((uint64_t)(m)->hdr_bytes + (m)->big_hdr + 9ull * (m)->nrec + 64ull)

// Output arena of one resident chunk: SFQ_NSTREAMS sub-ranges.
struct SfqArena {
    uint64_t off[SFQ_NSTREAMS];    // byte offset of each stream's region in the arena buffer
    uint32_t cap[SFQ_NSTREAMS];
    uint32_t size[SFQ_NSTREAMS];   // filled by the coders (0 = stream never created)
};

// Per-wave workspace of the coders: chunk w of the wave owns slice w of every table.
struct SfqWorkspace {
    uint8_t  *gtab;  uint64_t gtab_stride;  uint32_t hbits;     // base-context tables
    uint32_t *qtab;  uint64_t qtab_words;   uint32_t cbits;     // quality-context tables (hashed); cbits = ENTRIES of one table
    uint32_t *pw;                                                // 256-symbol model pools
    uint32_t gen_ahead2;                                         // base decoder: prefetch the table line two bases ahead
    uint8_t  *rec_scratch;                                       // header encoder: per-chunk scratch in global memory (null: shared memory)
};

