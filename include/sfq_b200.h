/* sfq_b200.h - C ABI of the B200-native slimfastq hot path (libsfq_b200.so).
 *
 * The reference has no FFI; its hot path sits behind C++ objects that are driven once per record
 * from exactly two loops.  The entry points below replace those loops wholesale (a batch of
 * records in, all streams out), so a host program - the reference's own main(), or the CLI in
 * slimfastq_b200/csrc/sfq_cli.cpp - binds them where it used to run:
 *
 *   sfq_compress / sfq_compress_device     <- UsrSave::encode()              usrs.cpp:392-407
 *        RecSave::save  recs.cpp:277-372 | GenSave::save  gens.hpp:89-93, gens.cpp:138-159
 *        QltSave::save  qlts.hpp:82-90, qlts.cpp:74-136 | UsrSave::get_record usrs.cpp:303-390
 *        XFileSave::put/put_chr/put_str xfile.cpp:66-99 | RCoder::Encode/done coder.hpp:52-81
 *   sfq_decompress / sfq_decompress_device <- UsrLoad::decode()              usrs.cpp:539-574
 *        RecLoad::load  recs.cpp:374-461 | GenLoad::load  gens.hpp:110-114, gens.cpp:215-249
 *        QltLoad::load  qlts.hpp:108-116, qlts.cpp:163-234 | UsrLoad::update/save usrs.cpp:471-529
 *        XFileLoad::get/get_chr/get_str xfile.cpp:76-109 | RCoder::GetFreq/Decode coder.hpp:83-102
 *   sfq_last_error + non-zero return       <- croak() + exit(1)              config.cpp:54-68
 *   the container written/read             <- FilerSave::put / FilerLoad::get filer.hpp:70-97
 *
 * Every chunk of the container holds the byte streams the reference would write for that chunk as
 * a standalone file at the same level (rec, gen, qlt, gen.Ns, gen.Nn, rec.x, usr.x, usr.x.q,
 * usr.pfg, usr.pfq) plus the semantic keys of its info stream; see sfq_container.h.
 *
 * Plain pointers and sizes only.  All entry points return 0 on success; on failure they return a
 * non-zero SFQ_ERR_* code and sfq_last_error() gives the message (the reference's croak text where
 * one exists).  There is no CPU fallback: without a usable CUDA device sfq_create() fails.
 * A context is not thread-safe; use one per host thread / per GPU.
 */
#ifndef SFQ_B200_H
#define SFQ_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sfq_ctx sfq_ctx;

enum {
    SFQ_ERR_NONE = 0,
    SFQ_ERR_CUDA = 1,        /* no device / CUDA runtime failure                               */
    SFQ_ERR_ARG = 2,         /* bad argument                                                   */
    SFQ_ERR_FASTQ = 3,       /* input is not FASTQ the reference would accept (croak text)      */
    SFQ_ERR_UNSUPPORTED = 4, /* input the reference itself cannot represent (first header > 399, */
                             /* empty first base line) or a chunk of 4 GiB+ per plane             */
    SFQ_ERR_FORMAT = 5,      /* not a b200 chunked .sfq container / corrupt container           */
    SFQ_ERR_NOMEM = 6,       /* host or device memory                                          */
    SFQ_ERR_SPACE = 7        /* caller-provided output buffer too small                        */
};

/* Per-call measurements (device times from CUDA events on the library's stream). */
typedef struct sfq_stats {
    uint64_t in_bytes, out_bytes;
    uint64_t nchunks, nrecords, nbases, nquals;
    uint64_t stream_bytes;        /* sum of all range-coded stream bytes                      */
    uint32_t waves;               /* coder launches (resident-chunk waves)                    */
    uint32_t resident_chunks;     /* chunks resident per wave                                 */
    uint32_t kernel_launches;     /* kernels launched by the call                             */
    uint32_t retries;             /* reruns after a stream arena / hash table had to grow     */
    float ms_total;               /* whole call on the device timeline                        */
    float ms_h2d, ms_d2h;         /* host<->device copies (host-buffer entry points only)     */
    float ms_scan;                /* newline scan: count + prefix + fill                      */
    float ms_plan;                /* chunk bounds + per-chunk framing facts                   */
    float ms_clear;               /* zeroing the model tables                                 */
    float ms_code;                /* k_encode / k_decode (sum over waves)                     */
    float ms_pack;                /* k_blob_offsets + k_pack  |  k_out_offsets + k_assemble   */
    uint64_t workspace_bytes;     /* model tables resident per wave                           */
    float ms_gen, ms_qlt, ms_rec; /* the three coder kernels of a wave run concurrently; each   */
                                  /* one's own duration, summed over waves                      */
    uint64_t gen_stream_bytes;    /* bytes of the `gen` streams (bases)                         */
    uint64_t qlt_stream_bytes;    /* bytes of the `qlt` streams (qualities)                     */
} sfq_stats;

/* Create a context on CUDA device `device` (-1 = current).  Fails if no device is usable. */
int  sfq_create(sfq_ctx **ctx, int device);
void sfq_destroy(sfq_ctx *ctx);
const char *sfq_last_error(const sfq_ctx *ctx);
const char *sfq_version(void);          /* "2.04/6 b200" - user version / internal format, config.cpp:44-45 */

/* Upper bound on resident chunks per coder wave (0 = as many as device memory allows). */
int sfq_set_max_resident(sfq_ctx *ctx, uint32_t chunks);

/* NUMA node of the host the CUDA device hangs off (sysfs numa_node of its PCI function), -1 if unknown.  Pinned staging
 * buffers are placed by first touch: a caller that binds its thread to that node's cores before it allocates them keeps
 * the host<->device copies off the inter-socket link (slimfastq_b200.api.bind_to_device_node does exactly that). */
int sfq_device_numa_node(int device);

/* Give back the device and pinned memory the context keeps between calls (grow-only workspace sized for the largest
 * call so far - up to ~90 % of the device after a 10 GB call).  Results returned by earlier calls become invalid. */
int sfq_trim(sfq_ctx *ctx);

/* The chunk grid of a PART of a file (a segment of a pipe, the shard of one GPU).  Chunk c of a whole file
 * holds the records that start in [c*chunk_bytes, (c+1)*chunk_bytes); "chunks are partitioned by index
 * across the GPUs" means every part must form those very chunks.  A part must begin with the first record
 * that starts at or after a grid line k*chunk_bytes; `phase` = (global offset of that record) - k*chunk_bytes.
 * It applies to the next sfq_compress / sfq_compress_device call on the context only.  The blobs of the parts,
 * laid out back to back with one index, are then byte-identical to the one-call container. */
int sfq_set_chunk_phase(sfq_ctx *ctx, uint64_t phase);

/* Pinned host staging buffers (optional, but pageable memory halves the PCIe rate). */
void *sfq_host_alloc(size_t bytes);
void  sfq_host_free(void *p);

/* Compress a whole FASTQ buffer held in HOST memory.  level 1..4 (clamped like config.cpp:231-236),
 * chunk_bytes = target chunk size (0 = 1 MiB).  *out points to a context-owned pinned buffer that
 * stays valid until the next sfq_compress on this context (sfq_decompress returns into a buffer of
 * its own, so a container can be handed straight back to it). */
int sfq_compress(sfq_ctx *ctx, const uint8_t *fastq, size_t n, int level, uint64_t chunk_bytes,
                 const uint8_t **out, size_t *out_n);
/* Same with input and output resident in DEVICE memory (16-byte aligned). */
int sfq_compress_device(sfq_ctx *ctx, const void *d_fastq, size_t n, int level, uint64_t chunk_bytes,
                        void *d_out, size_t out_cap, size_t *out_n);
/* Worst-case container size sfq_compress_device may need for n input bytes. */
size_t sfq_compress_bound(size_t n, uint64_t chunk_bytes);

/* Decompress a chunked .sfq container. */
int sfq_decompress(sfq_ctx *ctx, const uint8_t *sfq, size_t n, const uint8_t **out, size_t *out_n);
int sfq_decompress_device(sfq_ctx *ctx, const void *d_sfq, size_t n, void *d_out, size_t out_cap, size_t *out_n);
/* Size of the FASTQ text a container (host pointer; only the 72-byte header is read) decodes to. */
int sfq_decompressed_size(const uint8_t *sfq, size_t n, uint64_t *out_n, int *level);

int sfq_get_stats(const sfq_ctx *ctx, sfq_stats *st);

/* ---- per-plane test hooks ---------------------------------------------------------------------
 * The three class pairs the reference drives per record - GenSave/GenLoad (gens.hpp:87-115), QltSave/QltLoad
 * (qlts.hpp:79-116), RecSave/RecLoad (recs.hpp:87-107) with the UsrSave/UsrLoad framing streams - one plane at a time,
 * so that a stream that differs from the reference's can be isolated from outside the library.
 *   sfq_encode_<plane>_chunks   runs ONLY that plane's kernels over the FASTQ buffer; the container's chunks carry its
 *                               streams (gen: gen gen.Ns gen.Nn | qlt: qlt | rec: rec rec.x usr.*) and sizes 0 elsewhere
 *   sfq_decode_<plane>_chunks   runs ONLY that plane's decoder over a complete container and returns the plane as lines,
 *                               one per record: base lines (exception lists applied; positions whose quality is '!' read
 *                               as the coded base, the "'!' means N" rule needing the other plane, gens.cpp:200-213),
 *                               quality lines, id lines
 * Results are returned like sfq_compress / sfq_decompress return theirs. */
int sfq_encode_gen_chunks(sfq_ctx *ctx, const uint8_t *fastq, size_t n, int level, uint64_t chunk_bytes, const uint8_t **out, size_t *out_n);
int sfq_encode_qlt_chunks(sfq_ctx *ctx, const uint8_t *fastq, size_t n, int level, uint64_t chunk_bytes, const uint8_t **out, size_t *out_n);
int sfq_encode_rec_chunks(sfq_ctx *ctx, const uint8_t *fastq, size_t n, int level, uint64_t chunk_bytes, const uint8_t **out, size_t *out_n);
int sfq_decode_gen_chunks(sfq_ctx *ctx, const uint8_t *sfq, size_t n, const uint8_t **out, size_t *out_n);
int sfq_decode_qlt_chunks(sfq_ctx *ctx, const uint8_t *sfq, size_t n, const uint8_t **out, size_t *out_n);
int sfq_decode_rec_chunks(sfq_ctx *ctx, const uint8_t *sfq, size_t n, const uint8_t **out, size_t *out_n);

/* ---- interchange with the reference's own file format -----------------------
 * The reference writes an 8 KiB-page WORM container (filer.cpp:41-303); one file = one set of streams,
 * i.e. exactly one chunk of this library's container.  These are host-side format conversions: no
 * coding happens here and no CUDA device is needed.
 *   sfq_export_reference   single-chunk container (compress with chunk_bytes >= file size) -> a file the
 *                          unmodified reference binary decodes (stands where FilerSave / Config::init
 *                          wrote the pages, filer.cpp:166-242, config.cpp:330-349,382-386)
 *   sfq_import_reference   reference-written file -> single-chunk container for sfq_decompress (stands
 *                          where FilerLoad / Config::load_info read them, filer.cpp:246-303, config.cpp:87-107).
 *                          The file must carry orig.size (absent only when the reference read stdin). */
int    sfq_is_reference_file(const uint8_t *p, size_t n);
size_t sfq_export_reference_bound(const uint8_t *sfq, size_t n);
int    sfq_export_reference(const uint8_t *sfq, size_t n, const char *orig_filename,
                            uint8_t *out, size_t out_cap, size_t *out_n);
size_t sfq_import_reference_bound(size_t n);
int    sfq_import_reference(const uint8_t *ref, size_t n, uint8_t *out, size_t out_cap, size_t *out_n);

/* ---- record boundaries in FASTQ text (host-only helpers of the streaming CLI and of N-GPU sharding) -----------
 * The reference finds records while it pages its input (usrs.cpp:96-122, 303-390); a caller that cuts the
 * input itself - segments of a pipe, shards for several GPUs - needs the same notion without coding anything.
 * A line that starts with '@' is a header iff the line two below starts with '+' (a quality line may begin
 * with '@', but it is then followed by a header and a base line, never by a '+' line two below).
 *   sfq_record_start_at_or_after   smallest record start >= pos (n if there is none)
 *   sfq_last_record_start          largest record start in [1, n) whose '+' line is inside the buffer; 0 if none:
 *                                  everything before it is whole records, the tail is carried to the next segment */
size_t sfq_record_start_at_or_after(const uint8_t *fastq, size_t n, size_t pos);
size_t sfq_last_record_start(const uint8_t *fastq, size_t n);
/* Where to cut a buffered part of a stream so that the NEXT part starts on the chunk grid (sfq_set_chunk_phase):
 * `global_off` = offset in the whole stream of fastq[0] (itself such a start, or 0).  Returns the cut (bytes of
 * this part; the rest is carried over) and the phase of the part that begins there; 0 if the buffer holds no
 * grid line followed by a verifiable record start yet (read more, or at end of input code everything). */
size_t sfq_stream_cut(const uint8_t *fastq, size_t n, uint64_t global_off, uint64_t chunk_bytes, uint64_t *next_phase);

#ifdef __cplusplus
}
#endif
#endif
