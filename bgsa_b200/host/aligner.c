/*
 * aligner.c -- `aligner -q <query> -d <database> -f <result>`: the reference's command line
 * (original/BGSA_CPU/main.c:36-106, banded/BGSA_CPU/main.c) and host pipeline
 * (original/BGSA_CPU/cal_cpu.c:121-476, thread.c) rebuilt around the GPU batch entry
 * bgsa_align_batch_submit/_wait.  Plain C; all CUDA is behind the C ABI of libbgsa_b200.so.
 *
 * Same inputs (one sequence per line, equal lengths, ACGTN), same outputs:
 *   <result>        per read bucket, per ref bucket (<= REF_BUCKET_COUNT queries), per device:
 *                   [query][that device's subjects] scores, int16 (int8 for banded)
 *   <result>.info   int nblocks, int ndev, int64 nqueries, then per block int64 count[ndev], int extra_count
 * (cal_cpu.c:247-250,350-352, thread.c:149-158) -- the reference's `convert -r` reads them unchanged.
 *
 * Reference options keep their meaning: -q -d -f, -N (host threads: accepted, unused), -k (banded
 * threshold).  Additive options (the reference bakes these into the generated align_core.c):
 *   -a myers | semiglobal | banded | bitpal | bitpal-nonpacked | bitpal-semiglobal     algorithm   (default myers)
 *   -M <match> -I <mismatch> -G <gap>                              BitPAl scores (default 2 -3 -5)
 *   -m 0|1                                                         Myers sign: 0 = -distance (default), 1 = +distance
 *   -g <n>                                                         number of GPUs (default 1): contiguous subject
 *                                                                  ranges per device like original/BGSA_AVX512/cal_mic.c:459-481
 */
#include <getopt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/types.h>
#include <unistd.h>

#include "../../include/bgsa_b200.h"

#define READ_BUCKET_SIZE 114857600LL /* original/BGSA_CPU/config.h:6  */
#define REF_BUCKET_COUNT 100         /* original/BGSA_CPU/config.h:12 */
#define MAX_GPUS 16
#define TILE 32

static double now_s(void) { /* timer.h:5-9 */
    struct timeval t;
    gettimeofday(&t, NULL);
    return t.tv_sec + t.tv_usec / 1000000.0;
}

static void die(const char *msg, const char *arg) {
    printf(msg, arg);
    printf("\n");
    exit(1);
}

static FILE *open_file(const char *name, const char *mode) { /* file.c:9-18 */
    FILE *fp = fopen(name, mode);
    if (!fp) die("Error - can't open or create file: %s", name);
    return fp;
}

static void print_help(void) {
    printf("\nUsage: ./aligner [options]\n\nCommandline options:\n\n");
    printf("  -q <arg>\n\t Query file (one sequence per line).\n\n");
    printf("  -d <arg>\n\t Database file (one sequence per line, equal lengths).\n\n");
    printf("  -f <arg>\n\t Alignment result file. \n\n");
    printf("  -k <arg>\n\t Filter threshold (banded). \n\n");
    printf("  -a <arg>\n\t Algorithm: myers | semiglobal | banded | bitpal | bitpal-nonpacked | bitpal-semiglobal. \n\n");
    printf("  -M/-I/-G <arg>\n\t BitPAl match / mismatch / gap scores (default 2 -3 -5). \n\n");
    printf("  -g <arg>\n\t Number of GPUs. \n\n");
    exit(1);
}

/* mapping_table: A,C,G,T,N -> 0..4, everything else 0 (global.c:6-15) */
static unsigned char map_char(unsigned char c) {
    switch (c) { case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'N': return 4; default: return 0; }
}

int main(int argc, char **argv) {
    const char *file_query = NULL, *file_database = NULL, *file_result = "data/result.txt"; /* main.c:39 (handle_args overrides main.c:75) */
    bgsa_params_t prm;
    bgsa_params_default(&prm, BGSA_MYERS_GLOBAL);
    int algo = BGSA_MYERS_GLOBAL, ngpu = 1, have_k = 0;
    int M = 2, I = -3, G = -5, sign = -1, threshold = 31, verbose = 0;
    int c;
    while ((c = getopt(argc, argv, "t:q:d:f:n:N:R:Dk:a:M:I:G:m:g:v")) != -1) {
        switch (c) {
            case 'q': file_query = optarg; break;
            case 'd': file_database = optarg; break;
            case 'f': file_result = optarg; break;
            case 'N': case 't': case 'n': case 'D': break; /* CPU / MIC thread options of the reference: no effect */
            case 'k': threshold = atoi(optarg); have_k = 1; break;
            case 'a':
                if (!strcmp(optarg, "myers")) algo = BGSA_MYERS_GLOBAL;
                else if (!strcmp(optarg, "semiglobal")) algo = BGSA_MYERS_SEMIGLOBAL;
                else if (!strcmp(optarg, "banded")) algo = BGSA_BANDED_MYERS;
                else if (!strcmp(optarg, "bitpal")) algo = BGSA_BITPAL_PACKED;
                else if (!strcmp(optarg, "bitpal-nonpacked")) algo = BGSA_BITPAL_NONPACKED;
                else if (!strcmp(optarg, "bitpal-semiglobal")) algo = BGSA_BITPAL_PACKED_SEMIGLOBAL;
                else print_help();
                break;
            case 'M': M = atoi(optarg); break;
            case 'I': I = atoi(optarg); break;
            case 'G': G = atoi(optarg); break;
            case 'm': sign = atoi(optarg) == 0 ? -1 : 1; break;
            case 'g': ngpu = atoi(optarg); break;
            case 'v': verbose = 1; break;
            default: print_help();
        }
    }
    (void)have_k;
    if (!file_query) { printf("Query file can't be empty.\n"); exit(1); }          /* main.c:91-94 */
    if (!file_database) { printf("Database file can't be empty. \n"); exit(1); }   /* main.c:96-99 */
    if (ngpu < 1 || ngpu > MAX_GPUS) die("Error - bad GPU count: %s", "-g");
    bgsa_params_default(&prm, algo);
    if (algo == BGSA_BITPAL_PACKED || algo == BGSA_BITPAL_NONPACKED || algo == BGSA_BITPAL_PACKED_SEMIGLOBAL) { prm.match = M; prm.mismatch = I; prm.gap = G; }
    prm.threshold = threshold;
    prm.myers_sign = sign;
    const int esize = bgsa_result_size(algo);

    double total_start = now_s(), read_total_time = 0, write_total_time = 0, cal_total_time = 0;
    int devcount = 0;
    double init_time = now_s(), alloc_time = 0;
    /* Driver start-up grows with the number of GPUs it has to bring up (seconds on an 8-GPU box): unless the user
     * chose the devices, expose only the -g we are going to use. */
    if (!getenv("CUDA_VISIBLE_DEVICES")) {
        char vis[8 * MAX_GPUS] = "";
        for (int g = 0; g < ngpu; g++) snprintf(vis + strlen(vis), sizeof(vis) - strlen(vis), g ? ",%d" : "%d", g);
        setenv("CUDA_VISIBLE_DEVICES", vis, 1);
    }
    if (bgsa_device_count(&devcount) != BGSA_OK || devcount < ngpu) die("Error - CUDA devices unavailable: %s", bgsa_last_error());
    if (bgsa_init_devices(ngpu) != BGSA_OK) die("Error - %s", bgsa_last_error());   /* all contexts at once, not one after the other */
    init_time = now_s() - init_time;

    char info_name[4096];
    snprintf(info_name, sizeof(info_name), "%s.info", file_result);                /* main.c:80-87 */
    mkdir("data", 0755);                                                           /* create_folder("data", 0755), cal_cpu.c:198 */
    FILE *fp_ref = open_file(file_query, "rb");
    FILE *fp_read = open_file(file_database, "rb");
    FILE *fp_result = open_file(file_result, "wb+");
    FILE *fp_info = open_file(info_name, "wb+");

    /* ---- queries: whole file, mapped to codes in place, '\n' kept (file.c:117-140) */
    struct stat sb;
    stat(file_query, &sb);
    int64_t ref_size = sb.st_size;
    char *ref = (char *)malloc((size_t)ref_size + 2);
    if (fread(ref, 1, (size_t)ref_size, fp_ref) != (size_t)ref_size) die("Error - can't read file: %s", file_query);
    int ref_len = 0;
    while (ref_len < ref_size && ref[ref_len] != '\n') ref_len++;
    if (ref_size == 0 || ref_len == 0) die("Error - empty query file: %s", file_query);
    if (ref[ref_size - 1] != '\n') ref[ref_size++] = '\n';
    int64_t ref_count = ref_size / (ref_len + 1);
    for (int64_t i = 0; i < ref_size; i++) if (ref[i] != '\n') ref[i] = (char)map_char((unsigned char)ref[i]);
    int ref_bucket_num = ref_count > REF_BUCKET_COUNT ? (int)((ref_count + REF_BUCKET_COUNT - 1) / REF_BUCKET_COUNT) : 1;
    int ref_bucket_count = ref_count > REF_BUCKET_COUNT ? REF_BUCKET_COUNT : (int)ref_count;   /* cal_cpu.c:210-216 */

    /* ---- database geometry (file.c:44-115): length from the first line, whole rows per bucket */
    stat(file_database, &sb);
    int64_t read_total_size = sb.st_size;
    int read_len = 0;
    {
        int ch;
        while ((ch = fgetc(fp_read)) != EOF && ch != '\n') read_len++;
        rewind(fp_read);
    }
    if (read_total_size == 0 || read_len == 0) die("Error - empty database file: %s", file_database);
    const int64_t stride = read_len + 1;
    const int64_t total_rows = (read_total_size + 1) / stride;       /* last newline optional (cal_cpu.c:241) */
    int64_t bucket_bytes = READ_BUCKET_SIZE;
    if (getenv("BGSA_READ_BUCKET_SIZE")) bucket_bytes = atoll(getenv("BGSA_READ_BUCKET_SIZE"));   /* test knob: many small buckets */
    int64_t rows_per_bucket = bucket_bytes / stride;
    if (rows_per_bucket < 1) rows_per_bucket = 1;
    if (rows_per_bucket > total_rows) rows_per_bucket = total_rows;
    const int read_bucket_num = (int)((total_rows + rows_per_bucket - 1) / rows_per_bucket);
    if (bgsa_supported(&prm, ref_len, read_len) != BGSA_OK) die("Error - %s", bgsa_last_error());

    { /* .info header (cal_cpu.c:247-250) */
        fwrite(&read_bucket_num, sizeof(int), 1, fp_info);
        fwrite(&ngpu, sizeof(int), 1, fp_info);
        fwrite(&ref_count, sizeof(int64_t), 1, fp_info);
        fflush(fp_info);
    }

    /* ---- pinned ping-pong buffers (read_seq_a/b, align_results_a/b; cal_cpu.c:219-267) */
    char *rows_buf[2];
    void *res_buf[2];
    alloc_time = now_s();
    for (int b = 0; b < 2; b++) {
        rows_buf[b] = (char *)bgsa_malloc_host((size_t)(rows_per_bucket * stride + 64));
        res_buf[b] = bgsa_malloc_host((size_t)esize * (size_t)ref_bucket_count * (size_t)rows_per_bucket + 64);
        if (!rows_buf[b] || !res_buf[b]) die("Error - %s", bgsa_last_error());
    }
    alloc_time = now_s() - alloc_time;

    int64_t total_subjects = 0, rows_done = 0;
    int64_t next_rows = rows_per_bucket < total_rows ? rows_per_bucket : total_rows;
    double t0 = now_s();
    size_t got = fread(rows_buf[0], 1, (size_t)(next_rows * stride), fp_read);
    if (got < (size_t)(next_rows * stride)) rows_buf[0][got] = '\n';   /* file without final newline (file.c:64-72) */
    read_total_time += now_s() - t0;

    /* Two (read bucket x ref bucket) items are in flight per device (slot = item & 1, the reference's a/b ping-pong,
     * thread.c:35-170): item k+1 is queued before item k is waited for, so the GPUs never drain between buckets; results
     * are written in item order. */
    int64_t item = 0, item_nq[2] = {0, 0}, item_rows[2] = {0, 0};
    int pending[2] = {0, 0};                 /* slot holds a submitted item whose results are not written yet */
#define FINISH_ITEM(k)                                                                                         \
    do {                                                                                                       \
        const int s_ = (int)((k) & 1);                                                                         \
        if (!pending[s_]) break;                                                                               \
        pending[s_] = 0;                                                                                       \
        double c0_ = now_s();                                                                                  \
        for (int g = 0; g < ngpu; g++)                                                                         \
            if (bgsa_align_batch_wait(g, s_) != BGSA_OK) die("Error - %s", bgsa_last_error());                 \
        cal_total_time += now_s() - c0_;                                                                       \
        double w0_ = now_s();                 /* output_task_cpu, thread.c:149-158 */                           \
        fwrite(res_buf[s_], (size_t)esize, (size_t)item_nq[s_] * (size_t)item_rows[s_], fp_result);            \
        fflush(fp_result);                                                                                     \
        write_total_time += now_s() - w0_;                                                                     \
    } while (0)

    for (int rb = 0; rb < read_bucket_num; rb++) {
        const int cur = rb & 1;
        const int64_t rows = next_rows;
        bgsa_seq_t seq = {read_len, rows * stride, rows, 0, 0, rows_buf[cur]};
        /* contiguous subject ranges per device, multiples of the tile, last device takes the rest
         * (dispatch_task, original/BGSA_AVX512/global.c:86-) */
        int64_t dev_first[MAX_GPUS], dev_count[MAX_GPUS];
        {
            int64_t per = (rows / ngpu) / TILE * TILE, off = 0;
            for (int g = 0; g < ngpu; g++) {
                dev_first[g] = off;
                dev_count[g] = (g == ngpu - 1) ? rows - off : per;
                off += dev_count[g];
            }
        }
        { /* per-block record (cal_cpu.c:350-352); extra_count is always 0: no SIMD padding rows */
            int extra = 0;
            fwrite(dev_count, sizeof(int64_t), (size_t)ngpu, fp_info);
            fwrite(&extra, sizeof(int), 1, fp_info);
            fflush(fp_info);
        }
        for (int fb = 0; fb < ref_bucket_num; fb++, item++) {
            const int ref_start = fb * ref_bucket_count;
            const int ref_end = (fb == ref_bucket_num - 1) ? (int)ref_count : (fb + 1) * ref_bucket_count;
            const int nq = ref_end - ref_start;
            const int slot = (int)(item & 1);
            FINISH_ITEM(item);                               /* item - 2: the slot and its result buffer must be free */
            char *res = (char *)res_buf[slot];
            item_nq[slot] = nq; item_rows[slot] = rows; pending[slot] = 1;
            double c0 = now_s();
            for (int g = 0; g < ngpu; g++) {   /* device-major result blocks (cal_mic.c:470-471) */
                int rc = bgsa_align_batch_submit(&prm, ref + (int64_t)ref_start * (ref_len + 1), nq, ref_len, &seq, dev_first[g],
                                                 dev_count[g], res + (size_t)esize * (size_t)dev_first[g] * (size_t)nq,
                                                 dev_count[g], g, slot);
                if (rc != BGSA_OK) die("Error - %s", bgsa_last_error());
            }
            cal_total_time += now_s() - c0;
            /* overlap: while the GPUs work on this bucket, read the next one (input_task_cpu, thread.c:35-123).  The
             * other row buffer may still be read by the previous bucket's last item: finish that one first. */
            if (fb == 0 && rb + 1 < read_bucket_num) {
                FINISH_ITEM(item + 1);                       /* item - 1 */
                double r0 = now_s();
                next_rows = total_rows - (rows_done + rows) < rows_per_bucket ? total_rows - (rows_done + rows) : rows_per_bucket;
                got = fread(rows_buf[cur ^ 1], 1, (size_t)(next_rows * stride), fp_read);
                if (got < (size_t)(next_rows * stride)) rows_buf[cur ^ 1][got] = '\n';
                read_total_time += now_s() - r0;
            }
        }
        rows_done += rows;
        total_subjects += rows;
    }
    /* drain: the (up to two) items still in flight, in order */
    FINISH_ITEM(item);          /* item - 2 */
    FINISH_ITEM(item + 1);      /* item - 1 */
#undef FINISH_ITEM
    fclose(fp_ref); fclose(fp_read); fclose(fp_result); fclose(fp_info);
    double free_time = now_s();
    for (int b = 0; b < 2; b++) { bgsa_free_host(rows_buf[b]); bgsa_free_host(res_buf[b]); }
    free(ref);
    double total_end = now_s();
    free_time = total_end - free_time;

    /* the reference's statistics block (cal_cpu.c:459-475); cal time includes H2D/D2H here */
    printf("score is %d, %d, %d\n", prm.match, prm.mismatch, prm.gap);
    printf("read_total_time  is %.2fs\n", read_total_time);
    printf("write_total_time is %.2fs\n", write_total_time);
    printf("mem_total_time is   %.2fs\n", 0.0);
    printf("\n");
    printf("query_len    is %d\n", ref_len);
    printf("query_count  is %ld\n", (long)ref_count);
    printf("subject_len   is %d\n", read_len);
    printf("subject_count is %ld\n\n", (long)total_subjects);
    printf("gpu_count    is %d\n", ngpu);
    printf("cal_total_times     is %.2fs\n", cal_total_time);
    printf("total time          is %.2fs\n", total_end - total_start);
    printf("cal GCUPS is %.2f\n", 1.0 * ref_len * ref_count * read_len * total_subjects / cal_total_time / 1000000000);
    printf("Total GCUPS is %.2f\n", 1.0 * ref_len * ref_count * read_len * total_subjects / (total_end - total_start) / 1000000000);
    printf("\n\n");
    if (verbose)   /* additive: where the wall time outside read/cal/write goes */
        printf("gpu_init_time is %.3fs\nhost_alloc_time is %.3fs\nhost_free_time is %.3fs\n", init_time, alloc_time, free_time);
    return 0;
}
