// The chunked .sfq container that replaces the reference's 8 KiB-page WORM file (filer.cpp) on this
// path.  Host-only definitions shared by the C-ABI library, the CLI and the test emulation.
//
//   file   = SfqFileHeader | blob[0] | blob[1] | ... | u64 blob_offset[nchunks] (index)
//   blob   = SfqBlobHeader | rec_first bytes | stream bytes in SFQ_S_* order, unpadded
//
// A blob carries exactly what the reference keeps for a standalone file: the semantic keys of its
// info stream (config.level, llen, usr.solid, usr.2id, gen.N_byte, num_records, rec.first,
// qlt.extra.hi) and the named range-coded streams (rec gen qlt gen.Ns gen.Nn rec.x usr.x usr.x.q usr.pfg usr.pfq usr.lrec
// usr.lgen usr.lqlt), byte-identical to the reference's.
// The first 16 bytes are the stamp the reference sniffs (config.cpp:295-304); the next 16 tell the
// two container kinds apart.  All integers little-endian.
#pragma once
#include <stdint.h>
#include <string.h>
#include "sfq_common.cuh"

#define SFQ_STAMP      "whoami=slimfastq"           /* 16 bytes */
#define SFQ_KIND       "\nformat=b200.c2\n"         /* 16 bytes */
#define SFQ_BLOB_MAGIC 0x43514653u                  /* "SFQC" */
#define SFQ_INTERNAL_VERSION 6                      /* config.cpp:44 */

// Blob flag: the chunk was imported from a reference-written file, whose info stream does not record the
// plane sizes - nbases / nquals / hdr_bytes / out_len are upper bounds (from orig.size), the decoder
// establishes the real ones from the usr.* streams.
#define SFQ_BLOB_IMPORTED 1u
// Blob flag: the header stream (`rec`) is in the layout of format versions below 5 (the file's `version` info key is
// below 5 or missing): decimal fields as gaps against the previous header's text, everything else as strings
// (RecLoad::load_pre5, recs.cpp:397-398, 463-510).  Only ever set on import; this library writes version 6.
#define SFQ_BLOB_PRE5 2u

#pragma pack(push, 1)
struct SfqFileHeader {
    char     stamp[16];
    char     kind[16];
    uint32_t version;        // SFQ_INTERNAL_VERSION
    uint32_t level;          // config.level
    uint64_t orig_size;      // orig.size
    uint64_t nchunks;
    uint64_t chunk_bytes;
    uint64_t index_off;      // file offset of the blob-offset index
    uint64_t out_size;       // bytes the container decodes to (sum of the blobs' out_len)
};                           // 80 bytes
struct SfqBlobHeader {
    uint32_t magic;
    uint32_t level;
    uint64_t text_len;       // FASTQ bytes of the chunk in the original input
    uint64_t out_len;        // FASTQ bytes the chunk decodes to (UsrLoad::save, usrs.cpp:512-529)
    uint32_t nrec;           // num_records
    uint32_t nbases, nquals, hdr_bytes;
    int32_t  llen;           // llen
    uint8_t  solid, two_id, n_byte, pad;      // pad: flags, SFQ_BLOB_IMPORTED
    uint32_t extra_hi;       // qlt.extra.hi
    uint32_t rec_first_len;  // rec.first follows the header
    uint32_t q_used, g_used; // distinct quality / base contexts touched (decoder table sizing hint; 0 = unknown)
    uint32_t nbig, big_bases, big_quals, big_hdr;   // oversized records (usr.lrec/lgen/lqlt): count and line bytes, see SfqChunkMeta
    uint32_t ssize[SFQ_NSTREAMS];
};                           // 132 bytes
#pragma pack(pop)

static inline void sfq_file_header_init(SfqFileHeader *h, int level, uint64_t orig, uint64_t nchunks,
                                        uint64_t chunk_bytes, uint64_t index_off, uint64_t out_size) {
    memcpy(h->stamp, SFQ_STAMP, 16);
    memcpy(h->kind, SFQ_KIND, 16);
    h->version = SFQ_INTERNAL_VERSION; h->level = (uint32_t)level; h->orig_size = orig;
    h->nchunks = nchunks; h->chunk_bytes = chunk_bytes; h->index_off = index_off; h->out_size = out_size;
}
static inline bool sfq_is_chunked_container(const uint8_t *p, size_t n) {
    return n >= sizeof(SfqFileHeader) && !memcmp(p, SFQ_STAMP, 16) && !memcmp(p + 16, SFQ_KIND, 16);
}
static inline uint64_t sfq_blob_size(const SfqBlobHeader *b) {
    uint64_t s = sizeof(SfqBlobHeader) + b->rec_first_len;
    for (int k = 0; k < SFQ_NSTREAMS; k++) s += b->ssize[k];
    return s;
}

// Framing checks of an untrusted container, shared by the library (csrc/sfq_abi.cu) and the test emulation so that both
// refuse the same files.  0 = fine; 1 = corrupt index; 2 = bad blob header; 3 = blob header whose counts cannot match.
static inline int sfq_index_check(const SfqFileHeader *fh, size_t n) {
    return (fh->nchunks == 0 || fh->nchunks > 0x7fffffffull || fh->index_off > n || n - fh->index_off < fh->nchunks * 8) ? 1 : 0;
}
static inline int sfq_blob_check(const SfqBlobHeader *b, uint64_t off, size_t n) {
    if (b->magic != SFQ_BLOB_MAGIC || off > n || sfq_blob_size(b) > n - off || b->level < 1 || b->level > 4 ||
        b->nrec == 0 || b->rec_first_len > 399)             // (rec_first_len 0: every record of the chunk is oversized)
        return 2;
    // a record prints at least "@h\nb\n+\nq\n": reject headers whose counts cannot match their out_len
    if (!(b->pad & SFQ_BLOB_IMPORTED) && (b->out_len < 6ull * b->nrec || b->nbig > b->nrec ||
                                          (uint64_t)b->nbases + b->nquals + b->hdr_bytes + b->big_bases + b->big_quals + b->big_hdr > b->out_len))
        return 3;
    return 0;
}
