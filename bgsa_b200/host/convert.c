/*
 * convert.c -- `convert -f <fasta> | -q <fastq> | -r <result> [-o <out>]`: the reference's helper tool
 * (original/BGSA_CPU/convert.c, banded/BGSA_CPU/convert.c) rewritten for this code base.  Plain C, no GPU.
 *
 *   -f  FASTA  -> one sequence per line (convert.c:33-98): header lines ('>' up to the end of the line) are
 *       dropped, the lines of a record are joined, records are separated by '\n', one final '\n'.
 *   -q  FASTQ  -> one sequence per line (convert.c:100-165): '@' opens a header, the first line after it is the
 *       sequence, everything up to the next '@' ('+' line, qualities) is skipped.  Like the reference, a '@'
 *       inside a quality string is taken for a new header -- same bytes out for the same bytes in.
 *   -r  result file + "<result>.info" -> one score per line in query-major order (convert.c:167-277).  The
 *       payload is: per read bucket, per reference bucket (<= 100 queries), per device: [query][subjects of
 *       that device]; the last device of a bucket drops its `extra_count` padding scores.
 * The reference needs one `convert` binary per score width (common_write_t is int16_t in original/, int8_t in
 * banded/, config.h:19 / :21).  Additive option here:  -b 1|2  bytes per score (default 2; use 1 for banded).
 * Default output name: convert_result.txt (original/BGSA_CPU/convert.c:298).
 */
#include <getopt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define REF_BUCKET_COUNT 100 /* original/BGSA_CPU/config.h:12 */
#define IO_BUFFER (10 << 20)

static FILE *must_open(const char *name, const char *mode, const char *what) {
    FILE *fp = fopen(name, mode);
    if (!fp) {
        if (what) printf("%s\n", what);
        else printf("Error - can't open or create file: %s\n", name);
        exit(1);
    }
    return fp;
}

/* FASTA and FASTQ share one scanner: `open_char` starts a header that runs to the end of its line; sequence
 * bytes are copied without their newlines; with `single_line` only the first line after a header is sequence. */
static void convert_records(const char *input, const char *output, char open_char, int single_line) {
    FILE *in = must_open(input, "rb", NULL), *out = must_open(output, "w+", NULL);
    char *ibuf = malloc(IO_BUFFER), *obuf = malloc(IO_BUFFER);
    if (!ibuf || !obuf) { printf("Error - out of memory\n"); exit(1); }
    enum { SEQUENCE, HEADER, SKIP } state = SEQUENCE;
    int seen_record = 0;
    size_t n;
    while ((n = fread(ibuf, 1, IO_BUFFER, in)) > 0) {
        size_t o = 0;
        for (size_t i = 0; i < n; i++) {
            const char c = ibuf[i];
            if (c == open_char) {
                if (seen_record) obuf[o++] = '\n';
                seen_record = 1;
                state = HEADER;
            } else if (state == HEADER) {
                if (c == '\n') state = SEQUENCE;
            } else if (state == SEQUENCE) {
                if (c != '\n') obuf[o++] = c;
                else if (single_line) state = SKIP;
            }
        }
        /* the reference writes the chunk with fputs(): output stops at the first NUL byte of a chunk */
        size_t len = 0;
        while (len < o && obuf[len] != '\0') len++;
        fwrite(obuf, 1, len, out);
    }
    fputc('\n', out);
    fclose(in);
    fclose(out);
    free(ibuf);
    free(obuf);
}

static void convert_result(const char *result, const char *output, int score_bytes) {
    char *info_name = malloc(strlen(result) + 6);
    strcpy(info_name, result);
    strcat(info_name, ".info");
    FILE *in = must_open(result, "rb", "Can't read result file");
    FILE *info = must_open(info_name, "rb", "Can't read result info file");
    FILE *out = must_open(output, "w+", "Can't create output file");
    int nblocks = 0, ndev = 0;
    int64_t nqueries = 0;
    if (fread(&nblocks, sizeof(int), 1, info) != 1 || fread(&ndev, sizeof(int), 1, info) != 1 ||
        fread(&nqueries, sizeof(int64_t), 1, info) != 1 || nblocks < 0 || ndev <= 0) {
        printf("Can't read result info file\n");
        exit(1);
    }
    int64_t *counts = calloc((size_t)nblocks * ndev + 1, sizeof(int64_t));
    int *extra = calloc((size_t)nblocks + 1, sizeof(int));
    int64_t max_count = 0;
    for (int b = 0; b < nblocks; b++) {
        if (fread(counts + (size_t)b * ndev, sizeof(int64_t), ndev, info) != (size_t)ndev || fread(&extra[b], sizeof(int), 1, info) != 1) {
            printf("Can't read result info file\n");
            exit(1);
        }
        for (int d = 0; d < ndev; d++)
            if (counts[(size_t)b * ndev + d] > max_count) max_count = counts[(size_t)b * ndev + d];
    }
    for (int b = 0; b < nblocks; b++)
        for (int d = 0; d < ndev; d++) printf("read_count[%d][%d] is %d\n", b, d, (int)counts[(size_t)b * ndev + d]);

    /* start of every (read bucket, ref bucket) section of the payload */
    const int64_t nrefb = (nqueries + REF_BUCKET_COUNT - 1) / REF_BUCKET_COUNT;
    int64_t *section = calloc((size_t)nblocks * (nrefb > 0 ? nrefb : 1) + 1, sizeof(int64_t));
    int64_t pos = 0;
    for (int b = 0; b < nblocks; b++) {
        int64_t subjects = 0;
        for (int d = 0; d < ndev; d++) subjects += counts[(size_t)b * ndev + d];
        for (int64_t r = 0; r < nrefb; r++) {
            const int64_t nq = r == nrefb - 1 ? nqueries - r * REF_BUCKET_COUNT : REF_BUCKET_COUNT;
            section[(size_t)b * nrefb + r] = pos;
            pos += nq * subjects * score_bytes;
        }
    }
    unsigned char *buf = malloc((size_t)(max_count > 0 ? max_count : 1) * score_bytes);
    char *text = malloc(IO_BUFFER + 64);
    size_t tlen = 0;
    for (int64_t q = 0; q < nqueries; q++) {
        const int64_t r = q / REF_BUCKET_COUNT, qi = q % REF_BUCKET_COUNT;
        const int64_t nq = r == nrefb - 1 ? nqueries - r * REF_BUCKET_COUNT : REF_BUCKET_COUNT;
        for (int b = 0; b < nblocks; b++) {
            int64_t dev_start = section[(size_t)b * nrefb + r];
            for (int d = 0; d < ndev; d++) {
                const int64_t cnt = counts[(size_t)b * ndev + d];
                fseeko(in, dev_start + qi * cnt * score_bytes, SEEK_SET);
                const int64_t got = (int64_t)fread(buf, score_bytes, (size_t)cnt, in);
                const int64_t keep = cnt - (d == ndev - 1 ? extra[b] : 0);
                for (int64_t k = 0; k < keep; k++) {
                    int v = 0;
                    if (k < got) v = score_bytes == 1 ? (int)(int8_t)buf[k] : (int)(int16_t)(buf[2 * k] | (buf[2 * k + 1] << 8));
                    tlen += (size_t)sprintf(text + tlen, "%d\n", v);
                    if (tlen >= IO_BUFFER) { fwrite(text, 1, tlen, out); tlen = 0; }
                }
                dev_start += nq * cnt * score_bytes;
            }
        }
    }
    fwrite(text, 1, tlen, out);
    fclose(in); fclose(info); fclose(out);
    free(info_name); free(counts); free(extra); free(section); free(buf); free(text);
}

static void print_help(void) {
    printf("\nUsage: ./convert [options]\n\n");
    printf("Commandline options:\n\n");
    printf("  %-30s\n", "-f <arg>");
    printf("\t Convert the FATSA file to needed format. \n\n");
    printf("  %-30s\n", "-q <arg>");
    printf("\t Convert the FATSQ file to needed format. \n\n");
    printf("  %-30s\n", "-r <arg>");
    printf("\t Convert the result file to readable format. \n\n");
    printf("  %-30s\n", "-o <arg>");
    printf("\t Output file. \n\n");
    printf("  %-30s\n", "-b <arg>");
    printf("\t Bytes per score in the result file: 2 (default) or 1 (banded). \n\n");
    printf("  %-30s\n", "-h");
    printf("\t Print help. \n\n");
    exit(1);
}

int main(int argc, char **argv) {
    const char *input = NULL, *output = "convert_result.txt";
    int type = 0, score_bytes = 2, c;
    if (argc == 1) print_help();
    while ((c = getopt(argc, argv, "f:q:r:o:b:h")) != -1) {
        switch (c) {
            case 'f': type = 0; input = optarg; break;
            case 'q': type = 1; input = optarg; break;
            case 'r': type = 2; input = optarg; break;
            case 'o': output = optarg; break;
            case 'b': score_bytes = atoi(optarg); break;
            default: print_help();
        }
    }
    if (!input) { printf("Input file can't be empty.\n"); exit(1); }
    if (score_bytes != 1 && score_bytes != 2) { printf("-b must be 1 or 2\n"); exit(1); }
    if (type == 0) convert_records(input, output, '>', 0);
    else if (type == 1) convert_records(input, output, '@', 1);
    else convert_result(input, output, score_bytes);
    return 0;
}
