/* qz_api.cpp -- the qatzip.h C ABI on top of the B200 chunk engine.
 *
 * Mirrors, entry point by entry point, the behaviour of the reference's API layer
 * (src/qatzip.c qzInit:630, qzSetupSession*:1118-1345, qzCompress*:1842-2097,
 * qzDecompress*:2422-2671, qzTeardownSession:2673, qzClose:2748, qzMaxCompressedLength:3022,
 * qz{Set,Get}Defaults*:2780-2940; src/qatzip_stream.c; src/qatzip_mem.c) -- same argument
 * meaning, same return codes, same in/out length convention -- while everything below it
 * (device discovery, instance grabbing, DMA buffers, submit/poll threads, software fallback)
 * is replaced by qz_engine.cu.  No CPU codec exists in this library: without a usable CUDA
 * device every data-path call fails with QZ_NOSW_NO_HW.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <time.h>
#include <mutex>
#include <condition_variable>
#include <deque>
#include <thread>
#include <vector>
#include "../../include/qatzip.h"
#include "../../include/qatzip_b200.h"
#include <atomic>
#include "qz_engine.h"
#include "qz_crc32.h"

#define QZB_NUM_BUFF 32                 /* reference src/qatzip_internal.h:65 (req_cnt_thrshold ceiling) */
#define QZB_FMT_INTERNAL_LZ4 4          /* internal data_fmt value for LZ4 frames (QzbFormat) */
#define QZB_FMT_INTERNAL_ZLIB 5         /* DEFLATE_ZLIB: zlib_format=1 sessions (reference src/qatzip_internal.h:251) */

/* ------------------------------------------------------------------ logging */
static QzLogLevel_T g_log_level = LOG_WARNING;
extern "C" QzLogLevel_T qzSetLogLevel(QzLogLevel_T level) { QzLogLevel_T old = g_log_level; g_log_level = level; return old; }
/* exported because the reference CLI calls it through its QZ_PRINT/QZ_ERROR macros
 * (reference include/qz_utils.h:119-129, src/qatzip_utils.c:212) */
extern "C" void logMessage(QzLogLevel_T level, const char *file, int line, const char *format, ...)
{
    if (level > g_log_level) return;
    FILE *out = (level == LOG_ERROR || level == LOG_WARNING) ? stderr : stdout;
    static const char *names[] = { "", "Fatal", "Error", "Warning", "Info", "Debug", "Test", "Memory" };
    if (level == LOG_ERROR || level == LOG_WARNING) {
        char tb[32]; time_t now = time(NULL); struct tm tmv; localtime_r(&now, &tmv);
        strftime(tb, sizeof tb, "%Y-%m-%d %H:%M:%S", &tmv);
        fprintf(out, "[%s] [%s] (%s:%d): ", names[level], tb, file, line);
    } else if (level != LOG_NONE) fprintf(out, "[%s]: ", names[level]);
    va_list ap; va_start(ap, format); vfprintf(out, format, ap); va_end(ap);
}
#define QZ_ERR(...) logMessage(LOG_ERROR, __FILE__, __LINE__, __VA_ARGS__)
#define QZ_DBG(...) logMessage(LOG_DEBUG1, __FILE__, __LINE__, __VA_ARGS__)

/* ------------------------------------------------------------------ session state */
struct QzbParams {
    QzHuffmanHdr_T huffman_hdr; QzDirection_T direction; int data_fmt;   /* QzbFormat numbering */
    unsigned int comp_lvl; unsigned char comp_algorithm; unsigned int max_forks; unsigned char sw_backup;
    unsigned int hw_buff_sz, strm_buff_sz, input_sz_thrshold, req_cnt_thrshold, wait_cnt_thrshold;
    QzPollingMode_T polling_mode; unsigned int is_sensitive_mode;
    unsigned char stop_decompression_stream_end, zlib_format;
};
/* asynchronous requests (qzCompress2 / qzDecompress2 with a callback): one completion thread per
 * session consumes them in submission order and runs the ordinary synchronous engine, so a session
 * is still only ever driven by one thread at a time (reference: ring + AsyncReqConsumeJob,
 * src/qatzip.c:3854; lock-free ring src/qatzip_utils.c:1673-1823) */
struct QzbAsyncReq { int decompress; const unsigned char *src; unsigned char *dest; qzAsyncCallbackFn cb; QzResult_T *res; };
struct QzbAsync {
    std::mutex m; std::condition_variable cv, idle;
    std::deque<QzbAsyncReq> q;
    std::thread worker;
    bool stop = false, busy = false;
};
struct QzbSess {
    QzbParams p;
    QzbEngine *engine;
    unsigned char end_of_stream;
    QzB200Stats_T stats;
    QzbAsync *async;
};

static std::mutex g_lock;                     /* reference src/qatzip.c:124 g_lock */
static std::mutex g_defaults_lock;            /* reference src/qatzip.c:117 g_sess_params_lock */
static int g_init_done = 0, g_init_rc = QZ_NONE;
/* reference src/qatzip.c:97-116 defaults */
static QzbParams g_defaults = {
    QZ_HUFF_HDR_DEFAULT, QZ_DIRECTION_DEFAULT, QZ_DEFLATE_GZIP_EXT, QZ_COMP_LEVEL_DEFAULT, QZ_COMP_ALGOL_DEFAULT,
    QZ_MAX_FORK_DEFAULT, QZ_SW_BACKUP_DEFAULT, QZ_HW_BUFF_SZ, QZ_STRM_BUFF_SZ_DEFAULT, QZ_COMP_THRESHOLD_DEFAULT,
    QZB_NUM_BUFF, QZ_WAIT_CNT_THRESHOLD_DEFAULT, QZ_PERIODICAL_POLLING, 0, 0, 0
};

static int fmt_to_internal(QzDataFormat_T f) { return (int)f; }   /* 4B/GZIP/GZIP_EXT/RAW share numbering with QzbFormat */

/* ------------------------------------------------------------------ parameter validation
 * (reference src/qatzip_utils.c:395-635) */
static int check_common(unsigned direction, unsigned char algo, unsigned char sw_backup, unsigned hw, unsigned strm,
                        unsigned thr, unsigned req)
{
    if (direction > QZ_DIR_BOTH) return QZ_PARAMS;
    if (algo != QZ_DEFLATE && algo != QZ_LZ4 && algo != QZ_LZ4s && algo != QZ_ZSTD) return QZ_PARAMS;
    if (sw_backup > 1) return QZ_PARAMS;
    if (hw < QZ_HW_BUFF_MIN_SZ || hw > QZ_HW_BUFF_MAX_SZ) return QZ_PARAMS;
    if (strm < QZ_STRM_BUFF_MIN_SZ || strm > QZ_STRM_BUFF_MAX_SZ) return QZ_PARAMS;
    if (thr < QZ_COMP_THRESHOLD_MINIMUM) return QZ_PARAMS;
    if (req < QZ_REQ_THRESHOLD_MINIMUM || req > QZB_NUM_BUFF) return QZ_PARAMS;
    if (hw & (hw - 1)) return QZ_PARAMS;
    return QZ_OK;
}
static int check_params_v1(const QzSessionParams_T *p)
{
    if ((unsigned)p->huffman_hdr > QZ_STATIC_HDR) return QZ_PARAMS;
    if (p->comp_lvl < QZ_DEFLATE_COMP_LVL_MINIMUM || p->comp_lvl > QZ_DEFLATE_COMP_LVL_MAXIMUM) return QZ_PARAMS;
    if (p->comp_algorithm != QZ_DEFLATE) return QZ_PARAMS;
    if ((unsigned)p->data_fmt > QZ_DEFLATE_RAW) return QZ_PARAMS;
    return check_common(p->direction, p->comp_algorithm, p->sw_backup, p->hw_buff_sz, p->strm_buff_sz, p->input_sz_thrshold, p->req_cnt_thrshold);
}
static int check_params_common_t(const QzSessionParamsCommon_T *c)
{
    return check_common(c->direction, c->comp_algorithm, c->sw_backup, c->hw_buff_sz, c->strm_buff_sz, c->input_sz_thrshold, c->req_cnt_thrshold);
}
static int check_params_deflate(const QzSessionParamsDeflate_T *p)
{
    if (check_params_common_t(&p->common_params) != QZ_OK) return QZ_PARAMS;
    if (p->common_params.comp_algorithm != QZ_DEFLATE) return QZ_PARAMS;
    if ((unsigned)p->huffman_hdr > QZ_STATIC_HDR) return QZ_PARAMS;
    if (p->common_params.comp_lvl < 1 || p->common_params.comp_lvl > QZ_DEFLATE_COMP_LVL_MAXIMUM_Gen3) return QZ_PARAMS;
    if ((unsigned)p->data_fmt > QZ_DEFLATE_RAW) return QZ_PARAMS;
    return QZ_OK;
}
static int check_params_lz4(const QzSessionParamsLZ4_T *p)
{
    if (check_params_common_t(&p->common_params) != QZ_OK) return QZ_PARAMS;
    if (p->common_params.comp_algorithm != QZ_LZ4) return QZ_PARAMS;
    if (p->common_params.comp_lvl < QZ_LZS_COMP_LVL_MINIMUM || p->common_params.comp_lvl > QZ_LZS_COMP_LVL_MAXIMUM) return QZ_PARAMS;
    return QZ_OK;
}

static void common_from_internal(QzSessionParamsCommon_T *c, const QzbParams *p, unsigned char algo)
{
    memset(c, 0, sizeof *c);
    c->direction = p->direction; c->comp_lvl = p->comp_lvl; c->comp_algorithm = algo; c->max_forks = p->max_forks;
    c->sw_backup = p->sw_backup; c->hw_buff_sz = p->hw_buff_sz; c->strm_buff_sz = p->strm_buff_sz;
    c->input_sz_thrshold = p->input_sz_thrshold; c->req_cnt_thrshold = p->req_cnt_thrshold;
    c->wait_cnt_thrshold = p->wait_cnt_thrshold; c->polling_mode = p->polling_mode; c->is_sensitive_mode = p->is_sensitive_mode;
}
static void internal_from_common(QzbParams *p, const QzSessionParamsCommon_T *c)
{
    p->direction = c->direction; p->comp_lvl = c->comp_lvl; p->comp_algorithm = c->comp_algorithm; p->max_forks = c->max_forks;
    p->sw_backup = c->sw_backup; p->hw_buff_sz = c->hw_buff_sz; p->strm_buff_sz = c->strm_buff_sz;
    p->input_sz_thrshold = c->input_sz_thrshold; p->req_cnt_thrshold = c->req_cnt_thrshold;
    p->wait_cnt_thrshold = c->wait_cnt_thrshold; p->polling_mode = c->polling_mode; p->is_sensitive_mode = c->is_sensitive_mode;
}

/* ------------------------------------------------------------------ defaults API */
extern "C" int qzGetDefaults(QzSessionParams_T *d)
{
    if (!d) return QZ_PARAMS;
    std::lock_guard<std::mutex> g(g_defaults_lock);
    memset(d, 0, sizeof *d);
    d->huffman_hdr = g_defaults.huffman_hdr; d->direction = g_defaults.direction;
    d->data_fmt = g_defaults.data_fmt <= QZ_DEFLATE_RAW ? (QzDataFormat_T)g_defaults.data_fmt : QZ_DEFLATE_GZIP_EXT;
    d->comp_lvl = g_defaults.comp_lvl; d->comp_algorithm = QZ_DEFLATE; d->max_forks = g_defaults.max_forks;
    d->sw_backup = g_defaults.sw_backup; d->hw_buff_sz = g_defaults.hw_buff_sz; d->strm_buff_sz = g_defaults.strm_buff_sz;
    d->input_sz_thrshold = g_defaults.input_sz_thrshold; d->req_cnt_thrshold = g_defaults.req_cnt_thrshold;
    d->wait_cnt_thrshold = g_defaults.wait_cnt_thrshold;
    return QZ_OK;
}
extern "C" int qzSetDefaults(QzSessionParams_T *d)
{
    if (!d || check_params_v1(d) != QZ_OK) return QZ_PARAMS;
    std::lock_guard<std::mutex> g(g_defaults_lock);
    g_defaults.huffman_hdr = d->huffman_hdr; g_defaults.direction = d->direction; g_defaults.data_fmt = fmt_to_internal(d->data_fmt);
    g_defaults.comp_lvl = d->comp_lvl; g_defaults.comp_algorithm = d->comp_algorithm; g_defaults.max_forks = d->max_forks;
    g_defaults.sw_backup = d->sw_backup; g_defaults.hw_buff_sz = d->hw_buff_sz; g_defaults.strm_buff_sz = d->strm_buff_sz;
    g_defaults.input_sz_thrshold = d->input_sz_thrshold; g_defaults.req_cnt_thrshold = d->req_cnt_thrshold;
    g_defaults.wait_cnt_thrshold = d->wait_cnt_thrshold;
    return QZ_OK;
}
extern "C" int qzGetDefaultsDeflate(QzSessionParamsDeflate_T *d)
{
    if (!d) return QZ_PARAMS;
    std::lock_guard<std::mutex> g(g_defaults_lock);
    common_from_internal(&d->common_params, &g_defaults, QZ_DEFLATE);
    d->huffman_hdr = g_defaults.huffman_hdr;
    d->data_fmt = g_defaults.data_fmt <= QZ_DEFLATE_RAW ? (QzDataFormat_T)g_defaults.data_fmt : QZ_DEFLATE_GZIP_EXT;
    return QZ_OK;
}
extern "C" int qzSetDefaultsDeflate(QzSessionParamsDeflate_T *d)
{
    if (!d || check_params_deflate(d) != QZ_OK) return QZ_PARAMS;
    std::lock_guard<std::mutex> g(g_defaults_lock);
    internal_from_common(&g_defaults, &d->common_params);
    g_defaults.huffman_hdr = d->huffman_hdr; g_defaults.data_fmt = fmt_to_internal(d->data_fmt);
    return QZ_OK;
}
extern "C" int qzGetDefaultsDeflateExt(QzSessionParamsDeflateExt_T *d)
{
    if (!d) return QZ_PARAMS;
    int rc = qzGetDefaultsDeflate(&d->deflate_params);
    std::lock_guard<std::mutex> g(g_defaults_lock);
    d->stop_decompression_stream_end = g_defaults.stop_decompression_stream_end; d->zlib_format = g_defaults.zlib_format;
    return rc;
}
extern "C" int qzSetDefaultsDeflateExt(QzSessionParamsDeflateExt_T *d)
{
    if (!d || check_params_deflate(&d->deflate_params) != QZ_OK) return QZ_PARAMS;
    if (d->zlib_format > 1) return QZ_PARAMS;
    int rc = qzSetDefaultsDeflate(&d->deflate_params);
    std::lock_guard<std::mutex> g(g_defaults_lock);
    g_defaults.stop_decompression_stream_end = d->stop_decompression_stream_end;
    /* reference src/qatzip_utils.c:725-728: zlib_format turns the internal format into DEFLATE_ZLIB */
    g_defaults.zlib_format = d->zlib_format; if (d->zlib_format) g_defaults.data_fmt = QZB_FMT_INTERNAL_ZLIB;
    return rc;
}
extern "C" int qzGetDefaultsLZ4(QzSessionParamsLZ4_T *d)
{
    if (!d) return QZ_PARAMS;
    std::lock_guard<std::mutex> g(g_defaults_lock);
    common_from_internal(&d->common_params, &g_defaults, QZ_LZ4);
    return QZ_OK;
}
extern "C" int qzSetDefaultsLZ4(QzSessionParamsLZ4_T *d)
{
    if (!d || check_params_lz4(d) != QZ_OK) return QZ_PARAMS;
    std::lock_guard<std::mutex> g(g_defaults_lock);
    internal_from_common(&g_defaults, &d->common_params);
    g_defaults.data_fmt = QZB_FMT_INTERNAL_LZ4;
    return QZ_OK;
}
extern "C" int qzGetDefaultsLZ4S(QzSessionParamsLZ4S_T *d)
{
    if (!d) return QZ_PARAMS;
    std::lock_guard<std::mutex> g(g_defaults_lock);
    common_from_internal(&d->common_params, &g_defaults, QZ_LZ4s);
    d->qzCallback = NULL; d->qzCallback_external = NULL; d->lz4s_mini_match = 3;
    return QZ_OK;
}
extern "C" int qzSetDefaultsLZ4S(QzSessionParamsLZ4S_T *d) { (void)d; return QZ_UNSUPPORTED_FMT; }

/* ------------------------------------------------------------------ init / sessions */
extern "C" int qzInit(QzSession_T *sess, unsigned char sw_backup)
{
    if (!sess || sw_backup > 3) return QZ_PARAMS;
    std::lock_guard<std::mutex> g(g_lock);
    if (g_init_done && g_init_rc == QZ_OK) return QZ_DUPLICATE;        /* reference src/qatzip.c:651-654 */
    g_init_done = 1;
    g_init_rc = qzb_runtime_devices() > 0 ? QZ_OK : QZ_NOSW_NO_HW;     /* no device and no software engine */
    if (g_init_rc != QZ_OK) { sess->hw_session_stat = g_init_rc; QZ_ERR("no usable CUDA device: qatzip_b200 has no software path\n"); }
    return g_init_rc;
}
static int ensure_init(QzSession_T *sess)
{
    int rc = qzInit(sess, 1);
    return (rc == QZ_OK || rc == QZ_DUPLICATE) ? QZ_OK : rc;
}

static int attach_session(QzSession_T *sess, const QzbParams *p)
{
    if (sess->internal != NULL) return QZ_DUPLICATE;                    /* reference src/qatzip.c:1143-1145 */
    int rc = ensure_init(sess);
    if (rc != QZ_OK) { sess->hw_session_stat = rc; return rc; }
    QzbSess *s = (QzbSess *)calloc(1, sizeof(QzbSess));
    if (!s) { sess->hw_session_stat = QZ_NOSW_LOW_MEM; return QZ_NOSW_LOW_MEM; }
    s->p = *p;
    sess->internal = s;
    sess->hw_session_stat = QZ_OK;
    sess->thd_sess_stat = QZ_OK;
    sess->total_in = 0; sess->total_out = 0;
    return QZ_OK;
}

extern "C" int qzSetupSession(QzSession_T *sess, QzSessionParams_T *params)
{
    if (!sess) return QZ_PARAMS;
    QzSessionParams_T tmp;
    if (!params) { qzGetDefaults(&tmp); params = &tmp; }
    if (check_params_v1(params) != QZ_OK) return QZ_PARAMS;
    QzbParams p;
    { std::lock_guard<std::mutex> g(g_defaults_lock); p = g_defaults; }
    p.huffman_hdr = params->huffman_hdr; p.direction = params->direction; p.data_fmt = fmt_to_internal(params->data_fmt);
    p.comp_lvl = params->comp_lvl; p.comp_algorithm = params->comp_algorithm; p.max_forks = params->max_forks;
    p.sw_backup = params->sw_backup; p.hw_buff_sz = params->hw_buff_sz; p.strm_buff_sz = params->strm_buff_sz;
    p.input_sz_thrshold = params->input_sz_thrshold; p.req_cnt_thrshold = params->req_cnt_thrshold; p.wait_cnt_thrshold = params->wait_cnt_thrshold;
    p.stop_decompression_stream_end = 0; p.zlib_format = 0;
    return attach_session(sess, &p);
}
extern "C" int qzSetupSessionDeflate(QzSession_T *sess, QzSessionParamsDeflate_T *params)
{
    if (!sess) return QZ_PARAMS;
    QzSessionParamsDeflate_T tmp;
    if (!params) { qzGetDefaultsDeflate(&tmp); params = &tmp; }
    if (check_params_deflate(params) != QZ_OK) return QZ_PARAMS;
    QzbParams p;
    { std::lock_guard<std::mutex> g(g_defaults_lock); p = g_defaults; }
    internal_from_common(&p, &params->common_params);
    p.huffman_hdr = params->huffman_hdr; p.data_fmt = fmt_to_internal(params->data_fmt);
    p.stop_decompression_stream_end = 0; p.zlib_format = 0;
    return attach_session(sess, &p);
}
extern "C" int qzSetupSessionDeflateExt(QzSession_T *sess, QzSessionParamsDeflateExt_T *params)
{
    if (!sess) return QZ_PARAMS;
    QzSessionParamsDeflateExt_T tmp;
    if (!params) { qzGetDefaultsDeflateExt(&tmp); params = &tmp; }
    if (check_params_deflate(&params->deflate_params) != QZ_OK) return QZ_PARAMS;
    if (params->zlib_format > 1) return QZ_PARAMS;
    QzbParams p;
    { std::lock_guard<std::mutex> g(g_defaults_lock); p = g_defaults; }
    internal_from_common(&p, &params->deflate_params.common_params);
    p.huffman_hdr = params->deflate_params.huffman_hdr; p.data_fmt = fmt_to_internal(params->deflate_params.data_fmt);
    p.stop_decompression_stream_end = params->stop_decompression_stream_end; p.zlib_format = params->zlib_format;
    if (params->zlib_format) p.data_fmt = QZB_FMT_INTERNAL_ZLIB;       /* reference src/qatzip_utils.c:725-728 */
    return attach_session(sess, &p);
}
extern "C" int qzSetupSessionLZ4(QzSession_T *sess, QzSessionParamsLZ4_T *params)
{
    if (!sess) return QZ_PARAMS;
    QzSessionParamsLZ4_T tmp;
    if (!params) { qzGetDefaultsLZ4(&tmp); params = &tmp; }
    if (check_params_lz4(params) != QZ_OK) return QZ_PARAMS;
    QzbParams p;
    { std::lock_guard<std::mutex> g(g_defaults_lock); p = g_defaults; }
    internal_from_common(&p, &params->common_params);
    p.data_fmt = QZB_FMT_INTERNAL_LZ4; p.stop_decompression_stream_end = 0; p.zlib_format = 0;
    return attach_session(sess, &p);
}
extern "C" int qzSetupSessionLZ4S(QzSession_T *sess, QzSessionParamsLZ4S_T *params)
{
    (void)params;
    if (!sess) return QZ_PARAMS;
    return QZ_UNSUPPORTED_FMT;          /* QAT-specific intermediate format: out of scope */
}

extern "C" int qzTeardownSession(QzSession_T *sess)
{
    if (!sess) return QZ_PARAMS;
    if (sess->internal) {
        QzbSess *s = (QzbSess *)sess->internal;
        if (s->async) {
            { std::unique_lock<std::mutex> lk(s->async->m); s->async->idle.wait(lk, [&] { return s->async->q.empty() && !s->async->busy; }); s->async->stop = true; }
            s->async->cv.notify_all();
            s->async->worker.join();
            delete s->async;
        }
        if (s->engine) qzb_engine_destroy(s->engine);
        free(s);
        sess->internal = NULL;
    }
    return QZ_OK;
}
static void stream_pool_release();
extern "C" int qzClose(QzSession_T *sess)
{
    if (!sess) return QZ_PARAMS;
    /* the device context is process-wide and released at unload; what qzClose undoes is qzInit, so that a later qzInit
     * answers QZ_OK again like the reference's (src/qatzip.c:1084-1116 stops the service qzInit started) */
    stream_pool_release();
    std::lock_guard<std::mutex> g(g_lock);
    g_init_done = 0; g_init_rc = QZ_NONE;
    return QZ_OK;
}
extern "C" int qzGetStatus(QzSession_T *sess, QzStatus_T *status)
{
    if (!sess || !status) return QZ_PARAMS;
    memset(status, 0, sizeof *status);
    int n = qzb_runtime_devices();
    status->qat_hw_count = (unsigned short)(n > 0 ? n : 0);
    status->qat_service_init = n > 0; status->qat_mem_drvr = n > 0; status->qat_instance_attach = n > 0;
    status->hw_session_status = sess->hw_session_stat;
    status->algo_hw[QZ_DEFLATE] = n > 0; status->algo_hw[QZ_LZ4] = n > 0;
    return QZ_OK;
}
extern "C" int qzGetDeflateEndOfStream(QzSession_T *sess, unsigned char *eos)
{
    if (!sess || !eos || !sess->internal) return QZ_PARAMS;
    *eos = ((QzbSess *)sess->internal)->end_of_stream;
    return QZ_OK;
}

/* lazy default session + engine, like the reference does at the top of every data call
 * (reference src/qatzip.c:1894-1912) */
static int ready_session(QzSession_T *sess, QzbSess **out)
{
    int rc = ensure_init(sess);
    if (rc != QZ_OK) return rc;
    if (!sess->internal || sess->hw_session_stat == QZ_NONE) {
        int fmt;
        { std::lock_guard<std::mutex> g(g_defaults_lock); fmt = g_defaults.data_fmt; }
        rc = (fmt == QZB_FMT_INTERNAL_LZ4) ? qzSetupSessionLZ4(sess, NULL)
           : (fmt == QZB_FMT_INTERNAL_ZLIB) ? qzSetupSessionDeflateExt(sess, NULL) : qzSetupSessionDeflate(sess, NULL);
        if (rc != QZ_OK && rc != QZ_DUPLICATE) return rc;
    }
    QzbSess *s = (QzbSess *)sess->internal;
    if (!s->engine) {
        /* sessions take the configured devices in turn as their primary one (the reference hands its instances out
         * interleaved across devices, src/qatzip.c:795-808, qzGrabInstance :363); with one device that is the default device */
        static std::atomic<unsigned> g_next_device{0};
        int devs[16];
        const int nd = qzb_runtime_device_list(devs, 16);
        const int primary = nd > 1 ? devs[g_next_device.fetch_add(1) % (unsigned)nd] : qzb_runtime_default_device();
        s->engine = qzb_engine_create(primary);
        if (!s->engine) { sess->hw_session_stat = QZ_NOSW_NO_INST_ATTACH; return QZ_NOSW_NO_INST_ATTACH; }
    }
    *out = s;
    return QZ_OK;
}

/* ------------------------------------------------------------------ compress */
extern "C" int qzCompressCrcExt(QzSession_T *sess, const unsigned char *src, unsigned int *src_len, unsigned char *dest,
                                unsigned int *dest_len, unsigned int last, unsigned long *crc, uint64_t *ext_rc)
{
    int rc; QzbSess *s = NULL;
    if (!sess || !src || !src_len || !dest || !dest_len || (last != 0 && last != 1)) { rc = QZ_PARAMS; goto err; }
    if (ext_rc) *ext_rc = 0;
    rc = ready_session(sess, &s);
    if (rc != QZ_OK) goto err;
    {
        QzbCompressCall c; QzbCompressOut o;
        memset(&c, 0, sizeof c);
        c.fmt = s->p.data_fmt; c.level = (int)s->p.comp_lvl; c.static_huffman = (s->p.huffman_hdr == QZ_STATIC_HDR);
        c.last = (int)last; c.chunk_sz = s->p.hw_buff_sz;
        c.src = src; c.src_len = *src_len; c.dst = dest; c.dst_cap = *dest_len;
        c.src_pinned = qzb_pinned_contains(src, *src_len); c.dst_pinned = qzb_pinned_contains(dest, *dest_len);
        c.want_crc = crc != NULL; c.crc_in = crc ? (uint32_t)*crc : 0;
        rc = qzb_engine_compress(s->engine, &c, &o);
        s->stats.kernel_ms = o.kernel_ms; s->stats.codec_ms = o.codec_ms; s->stats.codec_launches = o.codec_launches; s->stats.kernel_launches = o.kernel_launches; s->stats.units = o.nchunks;
        s->stats.h2d_ms = o.h2d_ms; s->stats.d2h_ms = o.d2h_ms;
        if (rc != QZ_OK && rc != QZ_BUF_ERROR) goto err;
        if (crc && c.fmt != QZB_FMT_INTERNAL_LZ4) *crc = o.crc;       /* reference src/qatzip.c:1707-1714 */
        *src_len = (unsigned int)o.consumed; *dest_len = (unsigned int)o.produced;
        sess->total_in += o.consumed; sess->total_out += o.produced;
        sess->thd_sess_stat = rc;
        return rc;
    }
err:
    if (src_len) *src_len = 0;
    if (dest_len) *dest_len = 0;
    return rc;
}
extern "C" int qzCompressCrc(QzSession_T *sess, const unsigned char *src, unsigned int *src_len, unsigned char *dest,
                             unsigned int *dest_len, unsigned int last, unsigned long *crc)
{ return qzCompressCrcExt(sess, src, src_len, dest, dest_len, last, crc, NULL); }
extern "C" int qzCompressExt(QzSession_T *sess, const unsigned char *src, unsigned int *src_len, unsigned char *dest,
                             unsigned int *dest_len, unsigned int last, uint64_t *ext_rc)
{ return qzCompressCrcExt(sess, src, src_len, dest, dest_len, last, NULL, ext_rc); }
extern "C" int qzCompress(QzSession_T *sess, const unsigned char *src, unsigned int *src_len, unsigned char *dest,
                          unsigned int *dest_len, unsigned int last)
{ return qzCompressCrcExt(sess, src, src_len, dest, dest_len, last, NULL, NULL); }

/* ------------------------------------------------------------------ decompress */
extern "C" int qzDecompressCrcExt(QzSession_T *sess, const unsigned char *src, unsigned int *src_len, unsigned char *dest,
                                  unsigned int *dest_len, unsigned long *crc, uint64_t *ext_rc)
{
    (void)crc;   /* never written by the reference either (SURVEY.md section 8a notes) */
    int rc; QzbSess *s = NULL;
    if (!sess || !src || !src_len || !dest || !dest_len) { rc = QZ_PARAMS; goto err; }
    if (ext_rc) *ext_rc = 0;
    if (*src_len == 0) { *dest_len = 0; return QZ_OK; }                 /* reference src/qatzip.c:2465-2468 */
    rc = ready_session(sess, &s);
    if (rc != QZ_OK) goto err;
    {
        QzbDecompressCall c; QzbDecompressOut o;
        memset(&c, 0, sizeof c);
        c.fmt = s->p.data_fmt; c.chunk_sz = s->p.hw_buff_sz;
        c.src = src; c.src_len = *src_len; c.dst = dest; c.dst_cap = *dest_len;
        c.src_pinned = qzb_pinned_contains(src, *src_len); c.dst_pinned = qzb_pinned_contains(dest, *dest_len);
        c.stop_at_first = s->p.stop_decompression_stream_end;
        s->end_of_stream = 0;
        rc = qzb_engine_decompress(s->engine, &c, &o);
        s->stats.kernel_ms = o.kernel_ms; s->stats.codec_ms = o.kernel_ms; s->stats.codec_launches = o.kernel_launches; s->stats.kernel_launches = o.kernel_launches; s->stats.units = o.nmembers;
        /* an error behind members that did decode: their lengths are reported together with the code, like the reference does
         * (src/qatzip.c:2640-2650), so a stream caller can deliver them */
        if (rc != QZ_OK && rc != QZ_BUF_ERROR && o.consumed == 0) goto err;
        if (o.nmembers > 0) s->end_of_stream = 1;                        /* reference src/qatzip_utils.c:1534-1554 */
        *src_len = (unsigned int)o.consumed; *dest_len = (unsigned int)o.produced;
        sess->total_in += o.consumed; sess->total_out += o.produced;
        sess->thd_sess_stat = rc;
        return rc;
    }
err:
    if (src_len) *src_len = 0;
    if (dest_len) *dest_len = 0;
    return rc;
}
extern "C" int qzDecompressCrc(QzSession_T *sess, const unsigned char *src, unsigned int *src_len, unsigned char *dest,
                               unsigned int *dest_len, unsigned long *crc)
{ return qzDecompressCrcExt(sess, src, src_len, dest, dest_len, crc, NULL); }
extern "C" int qzDecompressExt(QzSession_T *sess, const unsigned char *src, unsigned int *src_len, unsigned char *dest,
                               unsigned int *dest_len, uint64_t *ext_rc)
{ return qzDecompressCrcExt(sess, src, src_len, dest, dest_len, NULL, ext_rc); }
extern "C" int qzDecompress(QzSession_T *sess, const unsigned char *src, unsigned int *src_len, unsigned char *dest,
                            unsigned int *dest_len)
{ return qzDecompressCrcExt(sess, src, src_len, dest, dest_len, NULL, NULL); }

/* async entry points.  NULL callback = the synchronous engine (reference src/qatzip.c:4122-4133);
 * with a callback the request is queued and the call returns QZ_OK at once; the callback later
 * receives the same QzResult_T with status / src_len (consumed) / dest_len (produced) filled in. */
static void async_worker(QzSession_T *sess, QzbAsync *a)
{
    for (;;) {
        QzbAsyncReq r;
        {
            std::unique_lock<std::mutex> lk(a->m);
            a->cv.wait(lk, [&] { return a->stop || !a->q.empty(); });
            if (a->q.empty()) return;
            r = a->q.front(); a->q.pop_front(); a->busy = true;
        }
        if (r.decompress) r.res->status = qzDecompressCrcExt(sess, r.src, &r.res->src_len, r.dest, &r.res->dest_len, NULL, &r.res->ext_rc);
        else r.res->status = qzCompressCrcExt(sess, r.src, &r.res->src_len, r.dest, &r.res->dest_len, 1, NULL, &r.res->ext_rc);
        r.cb(r.res);
        { std::lock_guard<std::mutex> lk(a->m); a->busy = false; }
        a->idle.notify_all();
    }
}
static int async_submit(QzSession_T *sess, int decompress, const unsigned char *src, unsigned char *dest, qzAsyncCallbackFn cb, QzResult_T *r)
{
    QzbSess *s = NULL;
    int rc = ready_session(sess, &s);
    if (rc != QZ_OK) { r->src_len = 0; r->dest_len = 0; return rc; }
    if (!s->async) { s->async = new QzbAsync(); s->async->worker = std::thread(async_worker, sess, s->async); }
    { std::lock_guard<std::mutex> lk(s->async->m); s->async->q.push_back(QzbAsyncReq{ decompress, src, dest, cb, r }); }
    s->async->cv.notify_one();
    return QZ_OK;
}
extern "C" int qzCompress2(QzSession_T *sess, const unsigned char *src, unsigned char *dest, qzAsyncCallbackFn cb, QzResult_T *r)
{
    if (!r) return QZ_PARAMS;
    if (!sess || !src || !dest) { r->src_len = 0; r->dest_len = 0; return QZ_PARAMS; }
    if (cb) return async_submit(sess, 0, src, dest, cb, r);
    r->status = qzCompressCrcExt(sess, src, &r->src_len, dest, &r->dest_len, 1, NULL, &r->ext_rc);
    return r->status;
}
extern "C" int qzDecompress2(QzSession_T *sess, const unsigned char *src, unsigned char *dest, qzAsyncCallbackFn cb, QzResult_T *r)
{
    if (!r) return QZ_PARAMS;
    if (!sess || !src || !dest) { r->src_len = 0; r->dest_len = 0; return QZ_PARAMS; }
    if (cb) return async_submit(sess, 1, src, dest, cb, r);
    r->status = qzDecompressCrcExt(sess, src, &r->src_len, dest, &r->dest_len, NULL, &r->ext_rc);
    return r->status;
}

/* ------------------------------------------------------------------ sizing
 * The reference's no-session bound (src/qatzip.c:3033-3044) also covers this build's worst
 * case (stored pieces: n + 11 bytes per piece + 32 bytes of framing per chunk << n/8), so it is
 * used for every session type. */
extern "C" unsigned int qzMaxCompressedLength(unsigned int src_sz, QzSession_T *sess)
{
    (void)sess;
    if (src_sz == 0) return QZ_COMPRESSED_SZ_OF_EMPTY_FILE;
    uint64_t out = ((uint64_t)src_sz * 9 + 7) / 8 + QZ_SKID_PAD_SZ + (24 + 8);
    if (out & 0xffffffff00000000ull) return 0;
    return (unsigned int)out;
}

/* ------------------------------------------------------------------ memory (reference src/qatzip_mem.c) */
extern "C" void *qzMalloc(size_t sz, int numa, int force_pinned)
{
    (void)numa;
    /* page-locked memory for PINNED_MEM requests and for anything large enough to be worth DMA-ing from directly;
     * small COMMON_MEM requests (stream bookkeeping, tests) come from the heap: cudaHostAlloc costs a system call and
     * locks pages (reference src/qatzip_mem.c:199-215: pinned first, heap as the COMMON_MEM fallback) */
    if (force_pinned == PINNED_MEM || sz >= (96u << 10)) {
        void *p = qzb_pinned_alloc(sz);
        if (p) return p;
        if (force_pinned == PINNED_MEM) return NULL;       /* reference src/qatzip_mem.c:211-215 */
    }
    return malloc(sz ? sz : 1);
}
extern "C" void qzFree(void *m)
{
    if (!m) return;
    if (!qzb_pinned_free(m)) free(m);
}
extern "C" int qzMemFindAddr(unsigned char *a) { return a ? qzb_pinned_contains(a, 1) : 0; }

/* ------------------------------------------------------------------ stream API (reference src/qatzip_stream.c) */
/* Compress streams are buffered three deep: while the caller fills one staging buffer, two worker threads of the stream
 * (each with an engine of its own, so their copies and kernels overlap) compress the two before it (reference: one synchronous
 * request per full buffer, src/qatzip_stream.c:514-560).  Jobs finish in any order and are taken over in stream order; each
 * job's CRC-32 is of its own bytes and is joined to the stream's by crc32_combine (the reference does the same per request,
 * src/qatzip.c:1707-1714). */
#define QZB_STREAM_WORKERS 2
struct QzbStreamJob {
    unsigned char *in = nullptr, *out = nullptr;     /* staging input (pinned), its compressed form */
    unsigned int in_cap = 0, out_cap = 0, in_len = 0, out_len = 0, last = 0, crc = 0;
    int rc = QZ_OK;                                  /* 1 = queued or running */
    int worker = 0;
};
struct QzbStreamWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    QzbSess *s = nullptr;
    QzbEngine *engine = nullptr;           /* worker 0: the session's; the others: their own, on the same device */
    bool own_engine = false;
    QzbStreamJob *job = nullptr;           /* queued or running; nullptr = idle */
    bool running = false, quit = false;
};
struct QzbStreamBuf {
    unsigned char *in_buf, *out_buf;
    unsigned int in_cap, out_cap;          /* staging capacity (may grow beyond strm_buff_sz on decode) */
    unsigned int in_off, out_off;          /* consumed prefix of in_buf / delivered prefix of out_buf */
    unsigned int flush_more;
    unsigned int finished;                 /* a call with last==1 has already been coded */
    unsigned int batched;                  /* compress: staging buffer already sized for batched engine calls */
    /* compress, asynchronous mode */
    QzbStreamWorker *worker[QZB_STREAM_WORKERS];
    QzbStreamJob *flight[QZB_STREAM_WORKERS];      /* jobs with the workers, oldest first */
    unsigned int nflight, last_queued, any_taken, nworkers;
    QzbStreamJob *spare;                   /* buffers of a finished job, reused for the next one */
};
/* What a stream sets up beyond its session -- the second worker's engine and the pinned staging buffers -- costs more than
 * compressing a gigabyte does, so qzEndStream parks them here for the next stream of the process (the reference sizes its
 * pinned pools once, in qzSetupSession: src/qatzip.c:1318-1400).  qzClose releases them. */
static std::mutex g_stream_pool_lock;
static std::vector<QzbEngine *> g_stream_engines;
struct QzbParkedBuf { unsigned char *p; unsigned int cap; };
static std::vector<QzbParkedBuf> g_stream_bufs;
static const size_t QZB_STREAM_POOL_BUFS = 12, QZB_STREAM_POOL_ENGINES = 4;
static QzbEngine *stream_engine_get(int device)
{
    {
        std::lock_guard<std::mutex> g(g_stream_pool_lock);
        for (size_t i = 0; i < g_stream_engines.size(); i++)
            if (qzb_engine_primary_device(g_stream_engines[i]) == device) {
                QzbEngine *e = g_stream_engines[i]; g_stream_engines.erase(g_stream_engines.begin() + (long)i); return e;
            }
    }
    return qzb_engine_create(device);
}
static void stream_engine_put(QzbEngine *e)
{
    {
        std::lock_guard<std::mutex> g(g_stream_pool_lock);
        if (g_stream_engines.size() < QZB_STREAM_POOL_ENGINES) { g_stream_engines.push_back(e); return; }
    }
    qzb_engine_destroy(e);
}
static unsigned char *stream_buf_get(unsigned int want, unsigned int *cap, bool exact = false)
{
    {
        std::lock_guard<std::mutex> g(g_stream_pool_lock);
        for (size_t i = 0; i < g_stream_bufs.size(); i++)
            if (exact ? g_stream_bufs[i].cap == want : (g_stream_bufs[i].cap >= want && g_stream_bufs[i].cap / 2 <= want)) {
                QzbParkedBuf b = g_stream_bufs[i]; g_stream_bufs.erase(g_stream_bufs.begin() + (long)i); *cap = b.cap; return b.p;
            }
    }
    *cap = want;
    return (unsigned char *)qzMalloc(want, QZ_AUTO_SELECT_NUMA_NODE, PINNED_MEM);
}
static void stream_buf_put(unsigned char *p, unsigned int cap)
{
    if (!p) return;
    if (cap >= (1u << 20) && qzb_pinned_contains(p, cap)) {
        std::lock_guard<std::mutex> g(g_stream_pool_lock);
        if (g_stream_bufs.size() < QZB_STREAM_POOL_BUFS) { g_stream_bufs.push_back({p, cap}); return; }
    }
    qzFree(p);
}
static void stream_pool_release()
{
    std::vector<QzbEngine *> es; std::vector<QzbParkedBuf> bs;
    { std::lock_guard<std::mutex> g(g_stream_pool_lock); es.swap(g_stream_engines); bs.swap(g_stream_bufs); }
    for (QzbEngine *e : es) qzb_engine_destroy(e);
    for (QzbParkedBuf &b : bs) qzFree(b.p);
}
static void stream_worker_main(QzbStreamWorker *w)
{
    std::unique_lock<std::mutex> lk(w->mu);
    for (;;) {
        w->cv.wait(lk, [&] { return w->quit || (w->job && !w->running && w->job->rc == 1); });
        if (w->quit) return;
        QzbStreamJob *j = w->job;
        w->running = true;
        lk.unlock();
        QzbCompressCall c; QzbCompressOut o;
        memset(&c, 0, sizeof c);
        c.fmt = w->s->p.data_fmt; c.level = (int)w->s->p.comp_lvl; c.static_huffman = (w->s->p.huffman_hdr == QZ_STATIC_HDR);
        c.last = (int)j->last; c.chunk_sz = w->s->p.hw_buff_sz;
        c.src = j->in; c.src_len = j->in_len; c.dst = j->out; c.dst_cap = j->out_cap;
        c.src_pinned = qzb_pinned_contains(j->in, j->in_len); c.dst_pinned = qzb_pinned_contains(j->out, j->out_cap);
        c.want_crc = 1; c.crc_in = 0;
        int rc = qzb_engine_compress(w->engine, &c, &o);
        if (rc == QZ_OK && o.consumed != j->in_len) rc = QZ_FAIL;
        lk.lock();
        j->out_len = (unsigned int)o.produced; j->crc = o.crc; j->rc = rc == QZ_OK ? QZ_OK : QZ_FAIL;
        w->running = false;
        w->cv.notify_all();
    }
}
static void stream_job_free(QzbStreamJob *j) { if (j) { stream_buf_put(j->in, j->in_cap); stream_buf_put(j->out, j->out_cap); delete j; } }
static int stream_init(QzSession_T *sess, QzStream_T *strm, QzbSess **sp)
{
    int rc = ready_session(sess, sp);
    if (rc != QZ_OK) return QZ_FAIL;
    if (strm->opaque) return QZ_OK;
    QzbStreamBuf *b = (QzbStreamBuf *)calloc(1, sizeof *b);
    if (!b) return QZ_FAIL;
    b->in_cap = (*sp)->p.strm_buff_sz; b->out_cap = (*sp)->p.strm_buff_sz;
    b->in_buf = (unsigned char *)qzMalloc(b->in_cap, QZ_AUTO_SELECT_NUMA_NODE, COMMON_MEM);
    b->out_buf = (unsigned char *)qzMalloc(b->out_cap, QZ_AUTO_SELECT_NUMA_NODE, COMMON_MEM);
    if (!b->in_buf || !b->out_buf) { qzFree(b->in_buf); qzFree(b->out_buf); free(b); return QZ_FAIL; }
    strm->opaque = b; strm->pending_in = 0; strm->pending_out = 0; strm->crc_32 = 0;
    return QZ_OK;
}
static unsigned int stream_copy_in(QzStream_T *strm, QzbStreamBuf *b, const unsigned char *in)
{
    unsigned int room = b->in_cap - b->in_off - strm->pending_in, n = strm->in_sz < room ? strm->in_sz : room;
    if (n) memcpy(b->in_buf + b->in_off + strm->pending_in, in, n);
    strm->pending_in += n; strm->in_sz -= n;
    return n;
}
static unsigned int stream_copy_out(QzStream_T *strm, QzbStreamBuf *b, unsigned char *out)
{
    unsigned int n = strm->pending_out < strm->out_sz ? strm->pending_out : strm->out_sz;
    if (n) memcpy(out, b->out_buf + b->out_off, n);
    strm->out_sz -= n; strm->pending_out -= n; b->out_off += n;
    if (strm->pending_out == 0) b->out_off = 0;
    return n;
}
static int grow(unsigned char **buf, unsigned int *cap, unsigned int keep_off, unsigned int keep_len, unsigned int want)
{
    unsigned char *n = (unsigned char *)qzMalloc(want, QZ_AUTO_SELECT_NUMA_NODE, want >= (1u << 20) ? PINNED_MEM : COMMON_MEM);
    if (!n) n = (unsigned char *)qzMalloc(want, QZ_AUTO_SELECT_NUMA_NODE, COMMON_MEM);
    if (!n) return QZ_FAIL;
    if (keep_len) memcpy(n, *buf + keep_off, keep_len);
    qzFree(*buf); *buf = n; *cap = want;
    return QZ_OK;
}

extern "C" int qzCompressStream(QzSession_T *sess, QzStream_T *strm, unsigned int last)
{
    if (!sess || !strm || (last != 0 && last != 1)) { if (strm) { strm->in_sz = 0; strm->out_sz = 0; } return QZ_PARAMS; }
    if (!strm->out || (!strm->in && strm->in_sz > 0)) { strm->in_sz = 0; strm->out_sz = 0; return QZ_PARAMS; }
    QzbSess *s = NULL;
    if (stream_init(sess, strm, &s) != QZ_OK) { strm->in_sz = 0; strm->out_sz = 0; return QZ_FAIL; }
    /* only these two formats stream: reference src/qatzip_stream.c:478-484 */
    if (s->p.data_fmt != QZ_DEFLATE_RAW && s->p.data_fmt != QZ_DEFLATE_GZIP_EXT) { strm->in_sz = 0; strm->out_sz = 0; return QZ_PARAMS; }
    QzbStreamBuf *b = (QzbStreamBuf *)strm->opaque;
    unsigned int consumed = 0, produced = 0; int rc = QZ_OK;
    /* The reference submits one synchronous engine request per full strm_buff_sz buffer (src/qatzip_stream.c:514-560).
     * A GPU launch costs about what 20 MiB of compression costs, so the staging buffer here is QZB200_STREAM_BATCH_KB
     * (default 8 MiB, a whole number of chunks) of pinned memory, and there are three of them: while the caller fills one,
     * the stream's two worker threads have their engines compress the others.  The stream is still cut into hw_buff_sz chunks, the
     * bytes are the same; output appears in larger steps.  QZB200_STREAM_BATCH_KB=0: the reference's cadence, synchronous. */
    if (!b->batched) {
        b->batched = 1;
        const char *ev = getenv("QZB200_STREAM_BATCH_KB");
        unsigned long long want = (ev && *ev ? strtoull(ev, NULL, 10) : 8192ull) << 10;
        const unsigned int hw = s->p.hw_buff_sz;
        if (want > (64ull << 20)) want = 64ull << 20;
        want = (want + hw - 1) / hw * hw;
        if (want > b->in_cap && strm->pending_in == 0 && strm->pending_out == 0) {
            unsigned int got = 0;
            unsigned char *ni = stream_buf_get((unsigned int)want, &got, true);     /* jobs are cut at the capacity: exactly this size */
            if (ni) { qzFree(b->in_buf); b->in_buf = ni; b->in_cap = got; b->in_off = 0; b->batched = 2; }
            const char *wv = getenv("QZB200_STREAM_WORKERS");        /* 1: one job in flight (A/B) */
            const long nw = wv && *wv ? strtol(wv, NULL, 10) : QZB_STREAM_WORKERS;
            b->nworkers = (unsigned int)(nw < 1 ? 1 : nw > QZB_STREAM_WORKERS ? QZB_STREAM_WORKERS : nw);
        }
    }
    if (b->batched == 2) {
        const unsigned int nworkers = b->nworkers;
        /* ---- asynchronous: up to QZB_STREAM_WORKERS jobs with the workers while this thread stages the next ---- */
        auto take_over = [&](bool wait) -> bool {        /* the oldest job's output becomes the pending output */
            if (!b->nflight || strm->pending_out) return false;
            QzbStreamJob *j = b->flight[0];
            QzbStreamWorker *w = b->worker[j->worker];
            {
                std::unique_lock<std::mutex> lk(w->mu);
                if (!wait && (w->running || j->rc == 1)) return false;
                w->cv.wait(lk, [&] { return !w->running && j->rc != 1; });
                w->job = nullptr;
            }
            for (unsigned int i = 1; i < b->nflight; i++) b->flight[i - 1] = b->flight[i];
            b->nflight--;
            if (b->spare) stream_job_free(b->spare);
            b->spare = j;
            if (j->rc != QZ_OK) { rc = QZ_FAIL; return false; }
            std::swap(b->out_buf, j->out); std::swap(b->out_cap, j->out_cap);
            strm->pending_out = j->out_len; b->out_off = 0;
            strm->crc_32 = b->any_taken ? qz_crc32_combine(strm->crc_32, j->crc, j->in_len) : j->crc;
            b->any_taken = 1;
            sess->total_in += j->in_len; sess->total_out += j->out_len;
            if (j->last) b->finished = 1;
            return true;
        };
        for (;;) {
            if (strm->pending_out) { produced += stream_copy_out(strm, b, strm->out + produced); if (strm->pending_out) break; }
            if (take_over(false)) continue;
            if (rc != QZ_OK) break;
            if (strm->in) consumed += stream_copy_in(strm, b, strm->in + consumed);
            const bool input_done = (strm->in_sz == 0);
            const bool want_last = last && input_done && !b->finished && !b->last_queued;
            if (strm->pending_in < b->in_cap && !want_last) {
                /* nothing to submit: at the end of the stream wait for what is still in flight */
                if (last && input_done && b->nflight) { if (take_over(true)) continue; if (rc != QZ_OK) break; }
                break;
            }
            /* the staged buffer goes to a worker: with all of them busy the oldest job must be done and its output taken over
             * first (if that output is still waiting for room in the caller's buffer, so does everything else) */
            if (b->nflight >= nworkers) { if (take_over(true)) continue; break; }
            int wi = -1;
            for (int i = 0; i < (int)nworkers && wi < 0; i++) {
                bool busy = false;
                for (unsigned int k = 0; k < b->nflight; k++) busy |= b->flight[k]->worker == i;
                if (!busy) wi = i;
            }
            if (!b->worker[wi]) {
                QzbStreamWorker *w = new QzbStreamWorker();
                w->s = s;
                if (wi == 0) w->engine = s->engine;
                else { w->engine = stream_engine_get(qzb_engine_primary_device(s->engine)); w->own_engine = true; }
                if (!w->engine) { delete w; rc = QZ_FAIL; break; }
                w->th = std::thread(stream_worker_main, w);
                b->worker[wi] = w;
            }
            QzbStreamJob *j = b->spare; b->spare = nullptr;
            const unsigned int need_out = qzMaxCompressedLength(b->in_cap, sess);
            if (!j) {
                j = new QzbStreamJob();
                j->in = stream_buf_get(b->in_cap, &j->in_cap, true);
                j->out = stream_buf_get(need_out, &j->out_cap);
                if (!j->in || !j->out) { stream_job_free(j); rc = QZ_FAIL; break; }
            }
            if (j->out_cap < need_out) {
                stream_buf_put(j->out, j->out_cap); j->out = stream_buf_get(need_out, &j->out_cap);
                if (!j->out) { stream_job_free(j); rc = QZ_FAIL; break; }
            }
            std::swap(b->in_buf, j->in); std::swap(b->in_cap, j->in_cap);        /* the job's old input buffer is the next one to fill */
            j->in_len = strm->pending_in; j->last = want_last ? 1u : 0u; j->out_len = 0; j->worker = wi;
            strm->pending_in = 0; b->in_off = 0;
            if (want_last) b->last_queued = 1;
            {
                std::lock_guard<std::mutex> g(b->worker[wi]->mu);
                j->rc = 1;                                   /* queued */
                b->worker[wi]->job = j;
            }
            b->flight[b->nflight++] = j;
            b->worker[wi]->cv.notify_all();
            if (want_last) { if (take_over(true)) continue; break; }
            if (input_done) break;
        }
        strm->in_sz = consumed; strm->out_sz = produced;
        return rc;
    }
    /* 1. hand over output left from an earlier call */
    if (strm->pending_out) {
        produced += stream_copy_out(strm, b, strm->out + produced);
        if (strm->pending_out) goto done;                 /* caller must make room first */
    }
    for (;;) {
        /* 2. batch input up to the staging capacity; only a full buffer or `last` triggers the engine */
        if (strm->in) consumed += stream_copy_in(strm, b, strm->in + consumed);
        const bool input_done = (strm->in_sz == 0);
        if (strm->pending_in < b->in_cap - b->in_off && !(last && input_done)) break;
        if (strm->pending_in == 0 && (!(last && input_done) || b->finished)) break;
        /* 3. one engine call over the staged bytes */
        unsigned int in_len = strm->pending_in, out_len = b->out_cap;
        unsigned int need = qzMaxCompressedLength(in_len ? in_len : 1, sess);
        if (need > b->out_cap) { if (grow(&b->out_buf, &b->out_cap, 0, 0, need) != QZ_OK) { rc = QZ_FAIL; break; } out_len = b->out_cap; }
        unsigned long crc = strm->crc_32;
        const unsigned int strm_last = (last && input_done) ? 1u : 0u;
        rc = qzCompressCrc(sess, b->in_buf + b->in_off, &in_len, b->out_buf, &out_len, strm_last, &crc);
        if (rc != QZ_OK) { rc = QZ_FAIL; break; }
        strm->crc_32 = (unsigned int)crc;
        if (strm_last) b->finished = 1;
        strm->pending_in -= in_len; b->in_off += in_len;
        if (strm->pending_in == 0) b->in_off = 0;
        strm->pending_out = out_len; b->out_off = 0;
        produced += stream_copy_out(strm, b, strm->out + produced);
        if (strm->pending_out) break;                     /* output full: resume on the next call */
        if (input_done) break;
    }
done:
    strm->in_sz = consumed; strm->out_sz = produced;
    return rc;
}

extern "C" int qzDecompressStream(QzSession_T *sess, QzStream_T *strm, unsigned int last)
{
    if (!sess || !strm || (last != 0 && last != 1)) { if (strm) { strm->in_sz = 0; strm->out_sz = 0; } return QZ_PARAMS; }
    if (!strm->in || !strm->out) { strm->in_sz = 0; strm->out_sz = 0; return QZ_PARAMS; }
    QzbSess *s = NULL;
    if (stream_init(sess, strm, &s) != QZ_OK) { strm->in_sz = 0; strm->out_sz = 0; return QZ_FAIL; }
    QzbStreamBuf *b = (QzbStreamBuf *)strm->opaque;
    unsigned int consumed = 0, produced = 0; int rc = QZ_OK;
    if (strm->pending_out) {
        produced += stream_copy_out(strm, b, strm->out + produced);
        if (strm->pending_out) goto done;
    }
    for (;;) {
        consumed += stream_copy_in(strm, b, strm->in + consumed);
        const bool input_done = (strm->in_sz == 0);
        const bool full = strm->pending_in >= b->in_cap - b->in_off;
        if (!full && !(last && input_done)) break;          /* batch more input */
        if (strm->pending_in == 0) break;
        unsigned int in_len = strm->pending_in, out_len = b->out_cap;
        rc = qzDecompress(sess, b->in_buf + b->in_off, &in_len, b->out_buf, &out_len);
        if (rc == QZ_OK || ((rc == QZ_BUF_ERROR || rc == QZ_DATA_ERROR) && in_len > 0)) {
            /* whole members were decoded; an incomplete tail stays staged */
            rc = QZ_OK;
            strm->pending_in -= in_len; b->in_off += in_len;
            if (strm->pending_in == 0) b->in_off = 0;
            else if (b->in_off) { memmove(b->in_buf, b->in_buf + b->in_off, strm->pending_in); b->in_off = 0; }
            strm->pending_out = out_len; b->out_off = 0;
            produced += stream_copy_out(strm, b, strm->out + produced);
            if (strm->pending_out) break;
            if (input_done && strm->pending_in == 0) break;
            if (input_done && !last && strm->pending_in < b->in_cap) break;     /* wait for more input */
            continue;                                                            /* more whole members may be staged */
        }
        if (in_len == 0 && (rc == QZ_BUF_ERROR || rc == QZ_DATA_ERROR || rc == QZ_PARAMS)) {
            /* no progress: the staged bytes do not hold one whole member, or its output does not
             * fit.  Unlike the reference (which hands this case to zlib's stateful inflate,
             * src/qatzip.c:2511-2530) the staging buffers grow until one member fits. */
            if (b->in_off) { memmove(b->in_buf, b->in_buf + b->in_off, strm->pending_in); b->in_off = 0; }
            if (rc == QZ_BUF_ERROR && b->out_cap < (1u << 30)) {
                if (grow(&b->out_buf, &b->out_cap, 0, 0, b->out_cap * 2) != QZ_OK) { rc = QZ_FAIL; break; }
                rc = QZ_OK; continue;
            }
            if (rc == QZ_DATA_ERROR && !(last && input_done) && b->in_cap < (1u << 30)) {
                if (grow(&b->in_buf, &b->in_cap, 0, strm->pending_in, b->in_cap * 2) != QZ_OK) { rc = QZ_FAIL; break; }
                rc = QZ_OK;
                if (input_done) break;
                continue;
            }
        }
        if (rc == QZ_OK) rc = QZ_FAIL;
        break;
    }
done:
    strm->in_sz = consumed; strm->out_sz = produced;
    return rc;
}

extern "C" int qzEndStream(QzSession_T *sess, QzStream_T *strm)
{
    if (!sess || !strm) return QZ_PARAMS;
    if (strm->opaque) {
        QzbStreamBuf *b = (QzbStreamBuf *)strm->opaque;
        for (int i = 0; i < QZB_STREAM_WORKERS; i++) {
            QzbStreamWorker *w = b->worker[i];
            if (!w) continue;
            { std::unique_lock<std::mutex> lk(w->mu); w->cv.wait(lk, [&] { return !w->running && !(w->job && w->job->rc == 1); }); w->quit = true; }
            w->cv.notify_all();
            w->th.join();
            if (w->own_engine) stream_engine_put(w->engine);
            delete w;
        }
        for (unsigned int i = 0; i < b->nflight; i++) stream_job_free(b->flight[i]);
        stream_job_free(b->spare);
        if (b->batched == 2) { stream_buf_put(b->in_buf, b->in_cap); stream_buf_put(b->out_buf, b->out_cap); }
        else { qzFree(b->in_buf); qzFree(b->out_buf); }
        free(b);
        strm->opaque = NULL;
    }
    strm->pending_in = 0; strm->pending_out = 0; strm->in_sz = 0; strm->out_sz = 0;
    return QZ_OK;
}

/* ------------------------------------------------------------------ device-resident extensions (include/qatzip_b200.h) */
extern "C" int qzb200CompressDevice(QzSession_T *sess, const void *d_src, uint64_t src_len, void *d_dest, uint64_t dest_cap,
                                    unsigned int last, uint64_t *consumed, uint64_t *produced, unsigned long *crc)
{
    QzbSess *s = NULL;
    if (!sess || !d_src || !d_dest || !consumed || !produced || (last != 0 && last != 1)) return QZ_PARAMS;
    int rc = ready_session(sess, &s);
    if (rc != QZ_OK) return rc;
    QzbCompressCall c; QzbCompressOut o;
    memset(&c, 0, sizeof c);
    c.fmt = s->p.data_fmt; c.level = (int)s->p.comp_lvl; c.static_huffman = (s->p.huffman_hdr == QZ_STATIC_HDR);
    c.last = (int)last; c.chunk_sz = s->p.hw_buff_sz;
    c.src = (const uint8_t *)d_src; c.src_len = src_len; c.dst = (uint8_t *)d_dest; c.dst_cap = dest_cap;
    c.src_device = 1; c.dst_device = 1; c.want_crc = crc != NULL; c.crc_in = crc ? (uint32_t)*crc : 0;
    rc = qzb_engine_compress(s->engine, &c, &o);
    s->stats.kernel_ms = o.kernel_ms; s->stats.codec_ms = o.codec_ms; s->stats.codec_launches = o.codec_launches; s->stats.kernel_launches = o.kernel_launches; s->stats.units = o.nchunks;
    *consumed = o.consumed; *produced = o.produced;
    if (crc && c.fmt != QZB_FMT_INTERNAL_LZ4 && (rc == QZ_OK || rc == QZ_BUF_ERROR)) *crc = o.crc;
    if (rc == QZ_OK || rc == QZ_BUF_ERROR) { sess->total_in += o.consumed; sess->total_out += o.produced; }
    return rc;
}
extern "C" int qzb200DecompressDevice(QzSession_T *sess, const void *d_src, const void *h_src_view, uint64_t src_len,
                                      void *d_dest, uint64_t dest_cap, uint64_t *consumed, uint64_t *produced)
{
    QzbSess *s = NULL;
    if (!sess || !d_src || !h_src_view || !d_dest || !consumed || !produced) return QZ_PARAMS;
    int rc = ready_session(sess, &s);
    if (rc != QZ_OK) return rc;
    QzbDecompressCall c; QzbDecompressOut o;
    memset(&c, 0, sizeof c);
    c.fmt = s->p.data_fmt; c.chunk_sz = s->p.hw_buff_sz;
    c.src = (const uint8_t *)d_src; c.src_host_view = (const uint8_t *)h_src_view; c.src_len = src_len;
    c.dst = (uint8_t *)d_dest; c.dst_cap = dest_cap; c.src_device = 1; c.dst_device = 1;
    c.stop_at_first = s->p.stop_decompression_stream_end;
    rc = qzb_engine_decompress(s->engine, &c, &o);
    s->stats.kernel_ms = o.kernel_ms; s->stats.codec_ms = o.kernel_ms; s->stats.codec_launches = o.kernel_launches; s->stats.kernel_launches = o.kernel_launches; s->stats.units = o.nmembers;
    *consumed = o.consumed; *produced = o.produced;
    if (rc == QZ_OK || rc == QZ_BUF_ERROR) { sess->total_in += o.consumed; sess->total_out += o.produced; }
    return rc;
}
extern "C" int qzb200GetStats(QzSession_T *sess, QzB200Stats_T *st)
{
    if (!sess || !st || !sess->internal) return QZ_PARAMS;
    QzbSess *s = (QzbSess *)sess->internal;
    QzbTuning t; qzb_get_tuning(&t);
    *st = s->stats; st->device = qzb_runtime_default_device(); st->devices = s->engine ? qzb_engine_device_count(s->engine) : 1; st->piece_log2 = t.piece_log2;
    st->group_blocks = (t.window && t.piece_log2 == 13 && s->p.hw_buff_sz % (8u << 13) == 0) ? 1 : 0;
    st->hash_bits = t.hash_bits;
    if (st->group_blocks) st->hash_bits = 11;        /* about 2600 entries per table: whatever the shared memory holds */
    return QZ_OK;
}
extern "C" void *qzb200DeviceAlloc(uint64_t n) { int d = qzb_runtime_default_device(); return d < 0 ? NULL : qzb_device_alloc(d, (size_t)n); }
extern "C" void qzb200DeviceFree(void *p) { int d = qzb_runtime_default_device(); if (d >= 0) qzb_device_free(d, p); }
extern "C" int qzb200CopyToDevice(void *d, const void *h, uint64_t n) { int dv = qzb_runtime_default_device(); return dv < 0 ? QZ_NOSW_NO_HW : qzb_device_copy(dv, d, h, (size_t)n, 1); }
extern "C" int qzb200CopyToHost(void *h, const void *d, uint64_t n) { int dv = qzb_runtime_default_device(); return dv < 0 ? QZ_NOSW_NO_HW : qzb_device_copy(dv, h, d, (size_t)n, 0); }
extern "C" int qzb200DeviceCount(void) { return qzb_runtime_devices(); }
extern "C" int qzb200DefaultDevice(void) { return qzb_runtime_default_device(); }

/* ------------------------------------------------------------------ stubs */
extern "C" int qzGetSoftwareComponentVersionList(QzSoftwareVersionInfo_T *a, unsigned int *n) { (void)a; (void)n; return QZ_FAIL; }
extern "C" int qzGetSoftwareComponentCount(unsigned int *n) { (void)n; return QZ_FAIL; }
#define NS(...) { return QZ_NOT_SUPPORTED; }
extern "C" int qzCompressCrc64(QzSession_T *, const unsigned char *, unsigned int *, unsigned char *, unsigned int *, unsigned int, uint64_t *) NS()
extern "C" int qzCompressCrc64Ext(QzSession_T *, const unsigned char *, unsigned int *, unsigned char *, unsigned int *, unsigned int, uint64_t *, uint64_t *) NS()
extern "C" int qzCompressWithMetadataExt(QzSession_T *, const unsigned char *, unsigned int *, unsigned char *, unsigned int *, unsigned int, uint64_t *, QzMetadataBlob_T, uint32_t, uint32_t) NS()
extern "C" int qzDecompressCrc64(QzSession_T *, const unsigned char *, unsigned int *, unsigned char *, unsigned int *, uint64_t *) NS()
extern "C" int qzDecompressCrc64Ext(QzSession_T *, const unsigned char *, unsigned int *, unsigned char *, unsigned int *, uint64_t *, uint64_t *) NS()
extern "C" int qzDecompressWithMetadataExt(QzSession_T *, const unsigned char *, unsigned int *, unsigned char *, unsigned int *, uint64_t *, QzMetadataBlob_T, uint32_t) NS()
extern "C" int qzAllocateMetadata(QzMetadataBlob_T *, size_t, uint32_t) NS()
extern "C" int qzFreeMetadata(QzMetadataBlob_T) NS()
extern "C" int qzGetSessionCrc64Config(QzSession_T *, QzCrc64Config_T *) NS()
extern "C" int qzGetSessionCrc32Config(QzSession_T *, QzCrc32Config_T *) NS()
extern "C" int qzSetSessionCrc64Config(QzSession_T *, QzCrc64Config_T *) NS()
extern "C" int qzSetSessionCrc32Config(QzSession_T *, QzCrc32Config_T *) NS()
extern "C" int qzMetadataBlockRead(uint32_t, QzMetadataBlob_T, uint32_t *, uint32_t *, uint32_t *, uint32_t *) NS()
extern "C" int qzMetadataBlockWrite(uint32_t, QzMetadataBlob_T, uint32_t *, uint32_t *, uint32_t *, uint32_t *) NS()
extern "C" int qzMetadataBlockGetCrc64(uint32_t, QzMetadataBlob_T, uint64_t *, uint64_t *) NS()
extern "C" int qzMetadataBlockGetCrc32(uint32_t, QzMetadataBlob_T, uint32_t *, uint32_t *) NS()
