/* stream_drive.c -- harness code: the piecemeal-submission loop of BASELINE config 5 (qzCompressStream fed in_step
 * bytes per call, the way reference test/main.c's stream tests feed it) written in C, so that a quarter of a
 * million 4 KiB calls are not timed through an interpreter.  The entry points are passed in as function
 * pointers taken from whichever library the caller loaded (product or reference build). */
#include <stdint.h>
#include <stddef.h>
#include <time.h>
#include "../include/qatzip.h"

typedef int (*stream_fn)(QzSession_T *, QzStream_T *, unsigned int);
typedef int (*end_fn)(QzSession_T *, QzStream_T *);

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

/* Feeds src[0..n) to fn in in_step-byte submissions (last = 1 on the final one), draining into out (reused, out_cap
 * bytes).  Returns the last return code; *seconds covers the loop only.  With sink != NULL the produced bytes are also
 * appended there (capacity sink_cap) so the caller can check the stream. */
int qzdrive_stream(void *fn_, void *end_, QzSession_T *sess, const unsigned char *src, size_t n, unsigned in_step,
                   unsigned char *out, unsigned out_cap, unsigned char *sink, size_t sink_cap,
                   uint64_t *out_bytes, uint64_t *calls, unsigned *crc, double *seconds)
{
    stream_fn fn = (stream_fn)fn_; end_fn end = (end_fn)end_;
    QzStream_T st = { 0 };
    size_t consumed = 0; uint64_t made = 0, ncalls = 0; int rc = QZ_OK;
    const double t0 = now_s();
    for (;;) {
        const size_t left = n - consumed;
        const unsigned step = left < in_step ? (unsigned)left : in_step;
        const unsigned last = (left - step == 0) ? 1u : 0u;
        st.in = (unsigned char *)src + consumed; st.in_sz = step; st.out = out; st.out_sz = out_cap;
        rc = fn(sess, &st, last);
        ncalls++;
        if (rc != QZ_OK) break;
        consumed += st.in_sz;
        if (sink && made + st.out_sz <= sink_cap) for (unsigned i = 0; i < st.out_sz; i++) sink[made + i] = out[i];
        made += st.out_sz;
        if (last && st.pending_in == 0 && st.pending_out == 0 && consumed == n) break;
    }
    *seconds = now_s() - t0;
    *out_bytes = made; *calls = ncalls; *crc = st.crc_32;
    end(sess, &st);
    return rc;
}

/* What the submission loop above costs with no compressor behind it: every in_step bytes are copied into a staging buffer
 * (the caller may reuse its own buffer after the call, so a stream API has to), and for every full staging buffer
 * out_per_stage bytes are copied out of a second one into out.  One thread, like the loop above; the seconds this takes
 * bound what any qzCompressStream can reach on this host. */
#include <string.h>
int qzdrive_copy_ceiling(const unsigned char *src, size_t n, unsigned in_step, unsigned char *stage_in, unsigned char *stage_out,
                         unsigned stage_cap, unsigned out_per_stage, unsigned char *out, double *seconds)
{
    size_t consumed = 0; unsigned fill = 0;
    if (out_per_stage > stage_cap) return -1;
    const double t0 = now_s();
    while (consumed < n) {
        const size_t left = n - consumed;
        unsigned step = left < in_step ? (unsigned)left : in_step;
        if (step > stage_cap - fill) step = stage_cap - fill;
        memcpy(stage_in + fill, src + consumed, step);
        consumed += step; fill += step;
        if (fill == stage_cap) { memcpy(out, stage_out, out_per_stage); fill = 0; }
    }
    *seconds = now_s() - t0;
    return 0;
}
