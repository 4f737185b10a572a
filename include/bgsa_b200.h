/*
 * include/bgsa_b200.h -- C ABI of libbgsa_b200.so, the B200 (sm_100a) implementation of BGSA's
 * one-query-versus-many-subjects alignment hot path.
 *
 * Plain C: pointers and sizes only, no CUDA or torch types.  Every entry point returns a
 * bgsa_status_t (0 = ok); nothing here falls back to the CPU -- if no CUDA device / kernel image
 * is available the call fails with BGSA_ERR_CUDA.
 *
 * What each entry point replaces in the reference (paths relative to sdu-hpcl/BGSA):
 *
 *   bgsa_align_batch            original/BGSA_CPU/cal.h:48  cpu_cal_align_score()  (cal_cpu.c:43-85)
 *                               + original/BGSA_CPU/global.h:24 cpu_handle_reads() (global.c:25-70),
 *                               and their sse_/avx_/mic_ twins (original/BGSA_AVX512/cal.h:54,
 *                               global.h:26) and the banded pair (banded/BGSA_CPU/cal.h:48,
 *                               global.h:24).  One call = one (read bucket x ref bucket x device)
 *                               iteration of cal_on_cpu() (cal_cpu.c:363-401) / mic_cal()
 *                               (original/BGSA_KNC/cal_mic.c:86-153).
 *   bgsa_pack_subjects_device   global.c:25-70 (Peq build) -- replaced by a 2-bit pack + transpose.
 *   bgsa_align_device           the generated kernels align_cpu/align_sse/align_avx/align_mic
 *                               (original/BGSA_CPU/align_core.h:8 and twins), batched.
 *   align_core.h (this dir)     the per-chunk kernel symbols themselves, for unmodified callers.
 *
 * Threading: every entry point may be called from any host thread; bgsa_last_error() is thread-local.  One job at a time
 * per (device, slot): a second bgsa_align_batch_submit on the same pair before bgsa_align_batch_wait returned would reuse
 * the pair's streams and buffers (the reference's a/b buffers have the same rule, thread.c:35-170).  Different devices and
 * the two slots of a device are independent.  bgsa_align_peq_chunk serialises its callers (it is called from an OpenMP team)
 * and has its own stream and buffers.  The device-resident entries (bgsa_align_device, bgsa_align_rows_device) keep their
 * query tables, work counters and packed scratch PER CALLER STREAM (up to 8 streams per device; a ninth recycles the
 * least recently used slot after draining its stream): calls on different streams never share state, calls on one stream
 * are ordered by the stream.
 *
 * Device pointers: d_packed must be 256-byte aligned and d_results aligned to the score size (BGSA_ERR_ARG otherwise --
 * a misaligned vector load would be a sticky context error).  d_rows may have ANY alignment; the kernels read it in whole
 * aligned 16-byte pieces, i.e. up to 15 bytes before its first and after its last byte inside the same 16-byte granules
 * (never written, never across a page the range does not touch).
 */
#ifndef BGSA_B200_H
#define BGSA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bgsa_status_t {
    BGSA_OK = 0,
    BGSA_ERR_ARG = 1,         /* null pointer, non-positive length, negative count ...            */
    BGSA_ERR_UNSUPPORTED = 2, /* scoring scheme / length / threshold has no kernel instance       */
    BGSA_ERR_CUDA = 3,        /* CUDA runtime error (message: bgsa_last_error())                  */
    BGSA_ERR_NOMEM = 4
} bgsa_status_t;

/* Algorithms on the path.  The numeric values are part of the ABI. */
typedef enum bgsa_algo_t {
    BGSA_MYERS_GLOBAL = 0,     /* generator -m, original/BGSA_CPU/align_core.c                    */
    BGSA_MYERS_SEMIGLOBAL = 1, /* generator -m -s, MyersGenerator.java:56-223                     */
    BGSA_BANDED_MYERS = 2,     /* generator -b, banded/BGSA_CPU/align_core.c, result int8         */
    BGSA_BITPAL_PACKED = 3,    /* generator -M -I -G, original/BGSA_AVX512/align_core.c           */
    BGSA_BITPAL_NONPACKED = 4, /* generator -t non-packed, BitPAlGenerator.java:1392-1701         */
    BGSA_BITPAL_PACKED_SEMIGLOBAL = 5 /* generator -s with -M/-I/-G: whole query inside the subject,
                                  BitPAlGenerator.java:77-80,112-114,289-308                      */
} bgsa_algo_t;

/* What the reference bakes into the generated align_core.c (align_core.c:13-17) plus -k. */
typedef struct bgsa_params_t {
    int32_t algo;        /* bgsa_algo_t                                                          */
    int32_t match;       /* BitPAl only (Myers/banded force 0,-1,-1: Main.java:253-257)          */
    int32_t mismatch;
    int32_t gap;
    int32_t threshold;   /* banded only: -k, default 31 (banded/BGSA_CPU/main.c:43)              */
    int32_t myers_sign;  /* Myers only: -1 => score = -distance (generator -m 0, the checked-in
                            kernels), +1 => +distance (-m 1, Main.java:127-141).  0 means -1.    */
} bgsa_params_t;

/* seq_t, bit-for-bit (original/BGSA_CPU/global.h:9-16). */
typedef struct bgsa_seq_t {
    int32_t len;
    int64_t size;
    int64_t count;
    int32_t extra_size;
    int32_t extra_count;
    char *content;       /* ASCII rows of stride len+1 (the '\n' is kept)                        */
} bgsa_seq_t;

/* ---- library / device ------------------------------------------------------------------ */
const char *bgsa_version(void);
const char *bgsa_last_error(void);                  /* thread-local, never NULL                  */
int bgsa_device_count(int *count);
/* Creates the CUDA contexts, streams and per-device state of devices 0 .. n_devices-1 IN PARALLEL (one host thread
 * per device).  Optional: every entry point initialises its device on first use, but context creation costs
 * 0.3-1 s per GPU and a multi-GPU caller (aligner -g 8) would otherwise pay it eight times in a row. */
int bgsa_init_devices(int n_devices);
void bgsa_params_default(bgsa_params_t *p, int algo); /* 2/-3/-5, threshold 31, sign -1          */
/* bytes per score: 1 for BGSA_BANDED_MYERS (common_write_t int8_t, banded/BGSA_CPU/config.h:21),
 * else 2 (int16_t, original/BGSA_CPU/config.h:19). */
int bgsa_result_size(int algo);
/* BGSA_OK iff (params, query_len, subject_len) has a kernel instance. */
int bgsa_supported(const bgsa_params_t *p, int query_len, int subject_len);

/* ---- host-buffer batch entry (the reference-facing call) ------------------------------- *
 * queries : code rows 0..4 of stride query_len+1, exactly what get_ref_from_file() leaves in
 *           ref_seq.content (file.c:117-140).
 * subjects: bgsa_seq_t with ASCII rows; subjects [first, first+count) are aligned.
 * results : [query][subject] row-major (cal_cpu.c:79-82), element (qi, si) at
 *           results[qi * result_stride + si], int16_t or int8_t (bgsa_result_size).
 * device  : CUDA ordinal.  Blocking; host<->device copies are inside the call.
 */
int bgsa_align_batch(const bgsa_params_t *p, const char *queries, int n_queries, int query_len,
                     const bgsa_seq_t *subjects, int64_t first, int64_t count,
                     void *results, int64_t result_stride, int device);

/* Asynchronous pair for double-buffered pipelines (thread.c:35-170 ping-pong): submit returns
 * as soon as the work is queued on the device's stream `slot` (0 or 1); wait blocks until the
 * results of that slot are in `results`.  Host buffers must stay valid until wait returns. */
int bgsa_align_batch_submit(const bgsa_params_t *p, const char *queries, int n_queries, int query_len,
                            const bgsa_seq_t *subjects, int64_t first, int64_t count,
                            void *results, int64_t result_stride, int device, int slot);
int bgsa_align_batch_wait(int device, int slot);
/* Pinned host memory for the buffers above (malloc_mem/free_mem, global.c:17-23). */
void *bgsa_malloc_host(size_t bytes);
void bgsa_free_host(void *p);
/* Pin buffers the caller already owns (the reference allocates its read/result buffers once per run with
 * malloc_mem, global.c:17-23, cal_cpu.c:258-267): page-locks [p, p+bytes) so that the copies of
 * bgsa_align_batch run at PCIe speed and overlap the kernels.  Unregister before freeing the buffer. */
int bgsa_host_register(void *p, size_t bytes);
int bgsa_host_unregister(void *p);
/* Binds the calling host thread to the CPUs of the NUMA node `device` hangs off (sysfs numa_node of its PCI
 * function), so that host buffers allocated afterwards are local to the GPU that will read them -- with one
 * process (or feeding thread) per GPU the 8 H2D streams of a box otherwise meet on one memory controller.
 * *numa_node receives the node, or -1 when the platform exposes none (then nothing is changed).  Linux only. */
int bgsa_bind_thread_to_device(int device, int *numa_node);

/* ---- device-resident entries (inputs already in HBM; used for kernel-only timing) ------ */
/* Size in bytes of the packed form of `count` subjects of `subject_len` bases. */
int64_t bgsa_packed_bytes(int subject_len, int64_t count);
/* d_rows: device pointer to ASCII rows (stride subject_len+1); d_packed: device buffer of
 * bgsa_packed_bytes().  p->algo selects the encoding the kernel of that algorithm consumes
 * (2-bit codes, or two bit-planes for BGSA_BANDED_MYERS).  stream: a cudaStream_t cast to
 * void* (NULL = default stream). */
int bgsa_pack_subjects_device(const bgsa_params_t *p, const void *d_rows, int subject_len, int64_t count,
                              void *d_packed, int device, void *stream);
/* The same encoding on the HOST cores (no GPU involved): `rows` and `packed` are host pointers, `packed` has
 * bgsa_packed_bytes() bytes and the layout bgsa_align_device consumes once copied to the device -- a database can be packed
 * once and kept (a quarter of the text's size).  Runs on the library's worker pool (BGSA_HOST_THREADS, default: all cores
 * / BGSA_HOST_GPUS).  bgsa_align_batch_submit uses the same code by itself when the PCIe copy of the ASCII rows, not the
 * kernel, would bound the batch and the host threads can encode faster than the link moves bytes (short reads;
 * BGSA_HOST_PACK=0/1 forces the choice) -- the GPU-era counterpart of the reference's OpenMP Peq build
 * (cpu_handle_reads, original/BGSA_CPU/global.c:25-70).  The N plane of a tile without 'N' is left unwritten, exactly as
 * the device pack kernels leave it (the per-tile flag says whether it is read). */
int bgsa_pack_subjects_host(const bgsa_params_t *p, const void *rows, int subject_len, int64_t count, void *packed);
/* threads of the worker pool and the encoder in use ("avx2" or "scalar"). */
int bgsa_host_pack_info(int *threads, char *isa, int isa_len);
/* h_queries: host code rows as above; d_results: device [query][count] scores. */
int bgsa_align_device(const bgsa_params_t *p, const char *h_queries, int n_queries, int query_len,
                      const void *d_packed, int subject_len, int64_t count,
                      void *d_results, int64_t result_stride, int device, void *stream);
/* d_rows: device pointer to ASCII rows (stride subject_len+1), as bgsa_pack_subjects_device takes them; scores to
 * d_results.  ONE kernel where the rows are short enough for a warp's shared-memory stage: banded Myers on rows up to ~950
 * bases (encodes every tile into shared memory and verifies it in place), Myers / BitPAl on queries up to 384 bases and rows
 * up to ~400 (thread per subject, match masks looked up by byte value: rows_kernel.cuh); otherwise packs into a
 * library-owned buffer and aligns.  bgsa_rows_kernel_name tells which. */
int bgsa_align_rows_device(const bgsa_params_t *p, const char *h_queries, int n_queries, int query_len,
                           const void *d_rows, int subject_len, int64_t count,
                           void *d_results, int64_t result_stride, int device, void *stream);
/* The front end of the last bgsa_align_batch[_submit] job on (device, slot): the share of its chunks that the host threads
 * packed (0 = every chunk crossed the link as ASCII, 1 = every chunk was packed on the host; -1 = decided chunk by chunk
 * from the measured state of the link: what a rank alone on its host does when neither extreme wins).  When several ranks
 * or devices share the host (torchrun's LOCAL_WORLD_SIZE, BGSA_HOST_GPUS, or several devices initialised by this process)
 * and the subjects are pinned, the share is tuned job by job on the measured throughput of the jobs themselves
 * (BGSA_HOST_PACK_NO_TUNING=1 keeps the static model's choice, BGSA_HOST_PACK_TUNING=1 tunes even alone;
 * BGSA_HOST_PACK=0/1/2 forces never / always / chunk by chunk). */
int bgsa_batch_front_end(int device, int slot, double *host_pack_share);
/* Name of what bgsa_align_rows_device runs for these parameters; *fused = 1 when that is ONE kernel fed with the ASCII rows
 * (no pack launch, no packed buffer), 0 when it is the pack kernel followed by bgsa_align_device's kernel. */
int bgsa_rows_kernel_name(const bgsa_params_t *p, int query_len, int subject_len, char *buf, int buflen, int *fused);
/* BitPAl scoring schemes.  The schemes listed at build time (make SCHEMES="2,-3,-5 ...") are template instances inside
 * the library.  Any other valid (match, mismatch, gap) -- gap < 0, match > mismatch, match >= 0, match - 2 gap <= 63
 * (non-packed: 24) after the common factor is divided out -- is instantiated on first use by NVRTC from the same kernel
 * headers (embedded in the library) and cached in $BGSA_JIT_CACHE (default ~/.cache/bgsa_b200): the run-time counterpart
 * of running the reference's generator for a new scheme (generator/.../Main.java:240-315).  bgsa_jit_precompile does that
 * compile ahead of time (no GPU needed); BGSA_OK at once for built-in schemes and for the Myers algorithms.  Without
 * libnvrtc such schemes report BGSA_ERR_UNSUPPORTED. */
int bgsa_jit_precompile(const bgsa_params_t *p, int query_len, int subject_len);
/* Number of kernel launches issued by this library since load (bench.py "gpu_launches"). */
int64_t bgsa_launch_count(void);
/* Name of the kernel instance bgsa_align_device would use, e.g. "bitpal_packed<2,-3,-5,K=5,L=1>". */
int bgsa_kernel_name(const bgsa_params_t *p, int query_len, int subject_len, char *buf, int buflen);

/* ---- per-chunk entry behind the reference's own kernel symbols ----------------------------- *
 * What align_cpu / align_sse / align_avx / align_mic (original/BGSA_CPU/align_core.h:8 and twins)
 * forward to, see include/align_core.h.  `peq` is the reference's match-mask block for
 * n_subjects subjects: [group][5][word_num][v_num] words of word_bytes bytes with usable_bits
 * cells per word (global.c:25-70); for BGSA_BANDED_MYERS the banded placement of
 * banded/BGSA_CPU/global.c:45-82 is assumed.  `results` receives n_subjects scores in subject
 * order.  Blocking, thread-safe (serialised). */
int bgsa_align_peq_chunk(const bgsa_params_t *p, const char *query, int query_len, const void *peq, int word_bytes,
                         int v_num, int usable_bits, int word_num, int subject_len, int64_t n_subjects, void *results,
                         int device);

/* ---- integer-pipe roofline probe ---------------------------------------------------------
 * Runs a LOP3/IADD3 throughput microbenchmark on `device` and reports lane-operations per
 * second on the INT32 ALU pipe (MEASURED_PEAKS.json has no integer figure). */
int bgsa_int_peak(int device, double *lane_ops_per_s, double *sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* BGSA_B200_H */
