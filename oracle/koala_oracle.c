/*
 * koala_oracle.c -- CPU restatement of the koala_b200 signal path (SPEC.md).  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (koala_b200/csrc -> libpv_koala_b200.so) never links, imports or calls it.
 *
 * PARITY STATUS: "parity unpinned" against the reference engine at sample level.  The reference
 * (Picovoice Koala v3.0.0) ships its per-frame path `pv_koala_process` (/root/reference/include/pv_koala.h:65-80)
 * only as a licence-activated closed binary (lib/linux/x86_64/libpv_koala.so) with no golden output vectors
 * (SURVEY.md F1-F3, F6).  What IS pinned, and what this file follows:
 *   - frame geometry: 256 int16 samples in, 256 int16 samples out per call     pv_koala.h:65-80  (frame_length=256 [bin])
 *   - fixed delay, output of call t = enhanced input of earlier calls          pv_koala.h:26-34, :92-100
 *   - reset == newly created, delayed samples are lost                          pv_koala.h:82-90
 *   - delay-trim / zero-flush loop of the file demo                             demo/c/koala_demo_file.c:466-521
 *   - behavioural tests (energy deviation < 0.02, bit-exact after reset)        binding/python/test_koala.py:71-129
 * Everything inside that contract (STFT geometry, features, network, rounding) is SPEC.md of this repository.
 *
 * Numerics: IEEE fp32, no fast-math, no FMA contraction (compile with -ffp-contract=off), accumulation over k
 * strictly sequential per output element, so the numpy restatement (oracle/numpy_oracle.py) can match it closely.
 * mode 0 ("fp32"): GEMM operands fp32 (weights are bf16-representable by construction).
 * mode 1 ("bf16"): every GEMM activation operand (features, encoder output, h) is rounded to bf16 (RNE) first;
 *                  accumulation, gates and the recurrent state stay fp32.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define KO_API __attribute__((visibility("default")))

enum { FRAME = 256, NFFT = 512, BINS = 256, MAXL = 8 };
static const float FEAT_POWER_SCALE = 9.31322574615478515625e-10f; /* 2^-30 */
static const float FEAT_EPS = 1e-6f, FEAT_GAIN = 0.125f, FEAT_BIAS = 0.25f;

typedef struct ko_model {
    int hidden, layers, bins;
    /* transposed fp32 copies: Wt[k][n] so the inner loop runs over outputs n */
    float *enc_wt, *enc_b;                 /* [bins][H], [H] */
    float *wih_t[MAXL], *whh_t[MAXL];      /* [H][3H] */
    float *bih[MAXL], *bhh[MAXL];          /* [3H] */
    float *dec_wt, *dec_b;                 /* [H][bins], [bins] */
    float window[NFFT];
    float tw_re[NFFT / 2], tw_im[NFFT / 2]; /* exp(-2 pi i k / 512) */
} ko_model_t;

/* ---------------------------------------------------------------- model file (format: koala_b200/spec.py) */
static uint32_t crc32_bytes(const uint8_t *p, size_t n) {
    static uint32_t table[256];
    static int init = 0;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int j = 0; j < 8; j++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = 1;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

static float bf16_bits_to_f32(uint16_t b) {
    uint32_t u = (uint32_t) b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static float bf16_round(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
    memcpy(&x, &u, 4);
    return x;
}

/* reads a [rows][cols] bf16 matrix and returns its fp32 transpose [cols][rows] */
static float *read_bf16_t(const uint8_t **p, int rows, int cols) {
    float *t = (float *) malloc(sizeof(float) * (size_t) rows * cols);
    const uint16_t *src = (const uint16_t *) *p;
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++) t[(size_t) c * rows + r] = bf16_bits_to_f32(src[(size_t) r * cols + c]);
    *p += 2 * (size_t) rows * cols;
    return t;
}

static float *read_f32(const uint8_t **p, int n) {
    float *t = (float *) malloc(sizeof(float) * n);
    memcpy(t, *p, sizeof(float) * n);
    *p += 4 * (size_t) n;
    return t;
}

KO_API void ko_model_free(ko_model_t *m) {
    if (!m) return;
    free(m->enc_wt); free(m->enc_b); free(m->dec_wt); free(m->dec_b);
    for (int l = 0; l < MAXL; l++) { free(m->wih_t[l]); free(m->whh_t[l]); free(m->bih[l]); free(m->bhh[l]); }
    free(m);
}

/* returns 0 on success; 1 io, 2 bad magic/geometry, 3 checksum/size */
KO_API int ko_model_load(const char *path, ko_model_t **out) {
    FILE *f = fopen(path, "rb");
    if (!f) return 1;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *blob = (uint8_t *) malloc(n > 0 ? n : 1);
    if (n < 48 || fread(blob, 1, n, f) != (size_t) n) { fclose(f); free(blob); return n < 48 ? 2 : 1; }
    fclose(f);
    if (memcmp(blob, "koala_b200\0\0", 12) != 0) { free(blob); return 2; }
    uint32_t hd[8];
    memcpy(hd, blob + 12, 32);
    if (hd[0] != 1 || hd[1] != NFFT || hd[2] != FRAME || hd[3] != BINS || hd[5] < 1 || hd[5] > MAXL || hd[6] != 1) {
        free(blob); return 2;
    }
    int H = (int) hd[4], L = (int) hd[5];
    size_t need = 44 + 2 * (size_t) H * BINS + 4 * (size_t) H + (size_t) L * (2 * 2 * 3 * (size_t) H * H + 2 * 4 * 3 * (size_t) H)
                  + 2 * (size_t) BINS * H + 4 * (size_t) BINS + 4;
    uint32_t crc;
    memcpy(&crc, blob + n - 4, 4);
    if ((size_t) n != need || crc != crc32_bytes(blob, n - 4)) { free(blob); return 3; }
    ko_model_t *m = (ko_model_t *) calloc(1, sizeof(*m));
    m->hidden = H; m->layers = L; m->bins = BINS;
    const uint8_t *p = blob + 44;
    m->enc_wt = read_bf16_t(&p, H, BINS);
    m->enc_b = read_f32(&p, H);
    for (int l = 0; l < L; l++) {
        m->wih_t[l] = read_bf16_t(&p, 3 * H, H);
        m->whh_t[l] = read_bf16_t(&p, 3 * H, H);
        m->bih[l] = read_f32(&p, 3 * H);
        m->bhh[l] = read_f32(&p, 3 * H);
    }
    m->dec_wt = read_bf16_t(&p, BINS, H);
    m->dec_b = read_f32(&p, BINS);
    free(blob);
    for (int i = 0; i < NFFT; i++) m->window[i] = (float) sin(M_PI * (double) i / NFFT);
    for (int k = 0; k < NFFT / 2; k++) {
        double a = -2.0 * M_PI * (double) k / NFFT;
        m->tw_re[k] = (float) cos(a);
        m->tw_im[k] = (float) sin(a);
    }
    *out = m;
    return 0;
}

KO_API int ko_model_hidden(const ko_model_t *m) { return m->hidden; }
KO_API int ko_model_layers(const ko_model_t *m) { return m->layers; }

/* ---------------------------------------------------------------- per-stream state */
typedef struct ko_stream {
    const ko_model_t *m;
    int mode;
    int16_t tail[FRAME];        /* previous input frame */
    float ola[FRAME];           /* second half of the previous synthesis frame (already windowed) */
    float *h;                   /* [layers][H] fp32 recurrent state */
    float last_mask[BINS];
    float last_feat[BINS];
} ko_stream_t;

KO_API void ko_stream_reset(ko_stream_t *s) {
    memset(s->tail, 0, sizeof(s->tail));
    memset(s->ola, 0, sizeof(s->ola));
    memset(s->h, 0, sizeof(float) * s->m->layers * s->m->hidden);
    memset(s->last_mask, 0, sizeof(s->last_mask));
    memset(s->last_feat, 0, sizeof(s->last_feat));
}

KO_API ko_stream_t *ko_stream_new(const ko_model_t *m, int mode) {
    ko_stream_t *s = (ko_stream_t *) calloc(1, sizeof(*s));
    s->m = m;
    s->mode = mode;
    s->h = (float *) calloc((size_t) m->layers * m->hidden, sizeof(float));
    ko_stream_reset(s);
    return s;
}

KO_API void ko_stream_free(ko_stream_t *s) {
    if (!s) return;
    free(s->h);
    free(s);
}

KO_API float *ko_stream_h(ko_stream_t *s) { return s->h; }
KO_API float *ko_stream_ola(ko_stream_t *s) { return s->ola; }
KO_API int16_t *ko_stream_tail(ko_stream_t *s) { return s->tail; }
KO_API const float *ko_stream_last_mask(const ko_stream_t *s) { return s->last_mask; }
KO_API const float *ko_stream_last_feat(const ko_stream_t *s) { return s->last_feat; }

/* ---------------------------------------------------------------- FFT (512-point complex, radix-2 DIT) */
static void fft512(const ko_model_t *m, float *re, float *im, int inverse) {
    /* bit reversal */
    for (int i = 0; i < NFFT; i++) {
        int j = 0;
        for (int b = 0; b < 9; b++) j |= ((i >> b) & 1) << (8 - b);
        if (j > i) {
            float t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int half = 1; half < NFFT; half <<= 1) {
        int step = NFFT / (2 * half); /* twiddle stride in units of W_512 */
        for (int base = 0; base < NFFT; base += 2 * half) {
            for (int j = 0; j < half; j++) {
                float wr = m->tw_re[j * step], wi = m->tw_im[j * step];
                if (inverse) wi = -wi;
                int a = base + j, b = a + half;
                float tr = re[b] * wr - im[b] * wi;
                float ti = re[b] * wi + im[b] * wr;
                re[b] = re[a] - tr; im[b] = im[a] - ti;
                re[a] = re[a] + tr; im[a] = im[a] + ti;
            }
        }
    }
}

/* analysis: frame = [tail | pcm] * window -> X[0..256]; packed spectrum spec[2k], spec[2k+1] = Re, Im X[k] for
 * k = 0..255 with spec[1] = Re X[256] (X[0] and X[256] are real); feat[k] for k = 0..255.  Updates tail. */
KO_API void ko_frontend(ko_stream_t *s, const int16_t *pcm, float *spec, float *feat) {
    const ko_model_t *m = s->m;
    float re[NFFT], im[NFFT];
    for (int n = 0; n < FRAME; n++) {
        re[n] = m->window[n] * (float) s->tail[n];
        re[n + FRAME] = m->window[n + FRAME] * (float) pcm[n];
    }
    memset(im, 0, sizeof(im));
    fft512(m, re, im, 0);
    for (int k = 0; k < BINS; k++) { spec[2 * k] = re[k]; spec[2 * k + 1] = im[k]; }
    spec[1] = re[256];
    for (int k = 0; k < BINS; k++) {
        float xi = (k == 0) ? 0.0f : im[k];
        float p = (re[k] * re[k] + xi * xi) * FEAT_POWER_SCALE;
        feat[k] = FEAT_GAIN * logf(p + FEAT_EPS) + FEAT_BIAS;
    }
    memcpy(s->tail, pcm, sizeof(s->tail));
    memcpy(s->last_feat, feat, sizeof(float) * BINS);
}

/* ---------------------------------------------------------------- mask estimator */
/* Y[s][n] = bias[n] + sum_k X[s][k] * Wt[k][n], k ascending, for S rows (S <= 8) */
static void gemm_rows(const float *X, int ldx, const float *Wt, const float *bias, float *Y, int ldy, int S, int K, int N) {
    enum { NB = 64 };
    for (int n0 = 0; n0 < N; n0 += NB) {
        int nb = N - n0 < NB ? N - n0 : NB;
        float acc[8][NB];
        for (int s = 0; s < S; s++)
            for (int n = 0; n < nb; n++) acc[s][n] = bias ? bias[n0 + n] : 0.0f;
        for (int k = 0; k < K; k++) {
            const float *w = Wt + (size_t) k * N + n0;
            for (int s = 0; s < S; s++) {
                float x = X[(size_t) s * ldx + k];
                for (int n = 0; n < nb; n++) acc[s][n] += x * w[n];
            }
        }
        for (int s = 0; s < S; s++)
            for (int n = 0; n < nb; n++) Y[(size_t) s * ldy + n0 + n] = acc[s][n];
    }
}

static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

static void round_rows(const float *src, float *dst, int n, int mode) {
    if (mode == 1) for (int i = 0; i < n; i++) dst[i] = bf16_round(src[i]);
    else memcpy(dst, src, sizeof(float) * n);
}

/* mask network for S streams at once (S <= 8): feat [S][256] -> mask [S][256]; updates each stream's h */
static void masknet_block(ko_stream_t **st, int S, const float *feat, float *mask, float *scratch) {
    const ko_model_t *m = st[0]->m;
    const int H = m->hidden, mode = st[0]->mode;
    float *xin = scratch;                 /* [S][max(H,256)] operand (rounded) */
    float *hin = xin + 8 * H;             /* [S][H] rounded recurrent operand */
    float *e = hin + 8 * H;               /* [S][H] layer input (fp32) */
    float *gi = e + 8 * H;                /* [S][3H] */
    float *gh = gi + 8 * 3 * H;           /* [S][3H] */
    for (int s = 0; s < S; s++) round_rows(feat + s * BINS, xin + s * H, BINS, mode);
    gemm_rows(xin, H, m->enc_wt, m->enc_b, e, H, S, BINS, H);
    for (int i = 0; i < S * H; i++) e[i] = e[i] > 0.0f ? e[i] : 0.0f;
    for (int l = 0; l < m->layers; l++) {
        for (int s = 0; s < S; s++) {
            round_rows(e + s * H, xin + s * H, H, mode);
            round_rows(st[s]->h + l * H, hin + s * H, H, mode);
        }
        gemm_rows(xin, H, m->wih_t[l], m->bih[l], gi, 3 * H, S, H, 3 * H);
        gemm_rows(hin, H, m->whh_t[l], m->bhh[l], gh, 3 * H, S, H, 3 * H);
        for (int s = 0; s < S; s++) {
            float *h = st[s]->h + l * H;
            const float *a = gi + s * 3 * H, *b = gh + s * 3 * H;
            for (int j = 0; j < H; j++) {
                float r = sigmoidf_(a[j] + b[j]);
                float z = sigmoidf_(a[H + j] + b[H + j]);
                float nn = tanhf(a[2 * H + j] + r * b[2 * H + j]);
                float hn = (1.0f - z) * nn + z * h[j];
                h[j] = hn;
                e[s * H + j] = hn;
            }
        }
    }
    for (int s = 0; s < S; s++) round_rows(e + s * H, xin + s * H, H, mode);
    gemm_rows(xin, H, m->dec_wt, m->dec_b, mask, BINS, S, H, BINS);
    for (int i = 0; i < S * BINS; i++) mask[i] = sigmoidf_(mask[i]);
    for (int s = 0; s < S; s++) memcpy(st[s]->last_mask, mask + s * BINS, sizeof(float) * BINS);
}

static size_t scratch_floats(int H) { return (size_t) 8 * H * 3 + (size_t) 8 * 3 * H * 2; }

KO_API void ko_masknet(ko_stream_t *s, const float *feat, float *mask) {
    float *scratch = (float *) malloc(sizeof(float) * scratch_floats(s->m->hidden));
    masknet_block(&s, 1, feat, mask, scratch);
    free(scratch);
}

/* ---------------------------------------------------------------- synthesis */
static inline int16_t round_sat(float v) {
    float r = rintf(v); /* round-half-even in the default rounding mode == cvt.rni */
    if (r > 32767.0f) r = 32767.0f;
    if (r < -32768.0f) r = -32768.0f;
    return (int16_t) r;
}

/* Y[k] = mask[k] X[k] (k = 0..255), Y[256] = mask[255] X[256]; y = irfft512(Y); s = window * y;
 * out[n] = round_sat(ola[n] + s[n]); ola'[n] = s[n + 256]. */
KO_API void ko_backend(ko_stream_t *s, const float *spec, const float *mask, int16_t *out) {
    const ko_model_t *m = s->m;
    float re[NFFT], im[NFFT];
    re[0] = mask[0] * spec[0]; im[0] = 0.0f;
    re[256] = mask[255] * spec[1]; im[256] = 0.0f;
    for (int k = 1; k < BINS; k++) {
        re[k] = mask[k] * spec[2 * k];
        im[k] = mask[k] * spec[2 * k + 1];
        re[NFFT - k] = re[k];
        im[NFFT - k] = -im[k];
    }
    fft512(m, re, im, 1);
    const float inv = 1.0f / NFFT;
    for (int n = 0; n < FRAME; n++) {
        float a = m->window[n] * (re[n] * inv);
        float b = m->window[n + FRAME] * (re[n + FRAME] * inv);
        out[n] = round_sat(s->ola[n] + a);
        s->ola[n] = b;
    }
}

KO_API void ko_stream_process(ko_stream_t *s, const int16_t *pcm, int16_t *out) {
    float spec[NFFT], feat[BINS], mask[BINS];
    ko_frontend(s, pcm, spec, feat);
    ko_masknet(s, feat, mask);
    ko_backend(s, spec, mask, out);
}

/* ---------------------------------------------------------------- batch of independent streams (CPU baseline) */
typedef struct ko_batch {
    const ko_model_t *m;
    int n, mode;
    ko_stream_t **st;
} ko_batch_t;

KO_API ko_batch_t *ko_batch_new(const ko_model_t *m, int n_streams, int mode) {
    ko_batch_t *b = (ko_batch_t *) calloc(1, sizeof(*b));
    b->m = m; b->n = n_streams; b->mode = mode;
    b->st = (ko_stream_t **) calloc(n_streams, sizeof(ko_stream_t *));
    for (int i = 0; i < n_streams; i++) b->st[i] = ko_stream_new(m, mode);
    return b;
}

KO_API void ko_batch_free(ko_batch_t *b) {
    if (!b) return;
    for (int i = 0; i < b->n; i++) ko_stream_free(b->st[i]);
    free(b->st);
    free(b);
}

KO_API void ko_batch_reset(ko_batch_t *b) { for (int i = 0; i < b->n; i++) ko_stream_reset(b->st[i]); }
KO_API ko_stream_t *ko_batch_stream(ko_batch_t *b, int i) { return b->st[i]; }

typedef struct { ko_batch_t *b; const int16_t *pcm; int16_t *out; int s0, s1, frames; size_t stride; } job_t;

static void *batch_worker(void *arg) {
    job_t *j = (job_t *) arg;
    const int H = j->b->m->hidden;
    float *scratch = (float *) malloc(sizeof(float) * scratch_floats(H));
    float spec[8][NFFT], feat[8 * BINS], mask[8 * BINS];
    for (int t = 0; t < j->frames; t++) {
        for (int s0 = j->s0; s0 < j->s1; s0 += 8) {
            int S = j->s1 - s0 < 8 ? j->s1 - s0 : 8;
            for (int s = 0; s < S; s++)
                ko_frontend(j->b->st[s0 + s], j->pcm + ((size_t) (s0 + s) * j->stride + (size_t) t * FRAME), spec[s], feat + s * BINS);
            masknet_block(j->b->st + s0, S, feat, mask, scratch);
            for (int s = 0; s < S; s++)
                ko_backend(j->b->st[s0 + s], spec[s], mask + s * BINS, j->out + ((size_t) (s0 + s) * j->stride + (size_t) t * FRAME));
        }
    }
    free(scratch);
    return NULL;
}

/* pcm/out: [n_streams][frames][256] int16 (stream-major); processes `frames` consecutive frames of every stream */
KO_API void ko_batch_process(ko_batch_t *b, const int16_t *pcm, int16_t *out, int frames, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > b->n) n_threads = b->n;
    pthread_t *th = (pthread_t *) malloc(sizeof(pthread_t) * n_threads);
    job_t *jobs = (job_t *) malloc(sizeof(job_t) * n_threads);
    int per = (b->n + n_threads - 1) / n_threads;
    per = (per + 7) / 8 * 8;
    int used = 0;
    for (int i = 0; i < n_threads; i++) {
        int s0 = i * per, s1 = s0 + per > b->n ? b->n : s0 + per;
        if (s0 >= s1) break;
        jobs[i] = (job_t){b, pcm, out, s0, s1, frames, (size_t) frames * FRAME};
        if (n_threads == 1) batch_worker(&jobs[i]);
        else pthread_create(&th[i], NULL, batch_worker, &jobs[i]);
        used++;
    }
    if (n_threads > 1) for (int i = 0; i < used; i++) pthread_join(th[i], NULL);
    free(th);
    free(jobs);
}
