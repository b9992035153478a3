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
 * mode 2 ("int8"): the fixed-point variant of SPEC.md section 6 (the reference engine's numeric style, SURVEY.md F4: int8
 *                  weights x int16 activations -> int32, table-driven gates): weights quantised per output row at load
 *                  time, int16 activations (features Q14, encoder output Q12, state Q15), int32 accumulation, 64-bit
 *                  requantisation multipliers, sigmoid / tanh from a 2049-entry table with linear interpolation.  Pure
 *                  integer arithmetic between the quantised features and the mask, so the CUDA path must match it BIT FOR
 *                  BIT there; analysis and synthesis stay fp32 as in the other modes.
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
    /* fixed-point variant (SPEC.md section 6), built on first use of mode 2 */
    struct ko_qmodel *q;
} ko_model_t;

/* ---------------------------------------------------------------- fixed-point model (SPEC.md section 6) */
enum { QF = 14, QE = 12, QH = 15, QP = 12, SIG_N = 2048 };
typedef struct ko_qmodel {
    int8_t *enc_q;                          /* [H][256] */
    int32_t *enc_m, *enc_b;                 /* [H] multiplier, bias (Q12) */
    int8_t *ih_q[MAXL], *hh_q[MAXL];        /* [3H][H], PyTorch row order r | z | n */
    int32_t *g_m[MAXL], *g_b[MAXL];         /* [4][H]: r, z, n_x, n_h */
    int8_t *dec_q;                          /* [256][H] */
    int32_t *dec_m, *dec_b;                 /* [256] */
    int16_t sig[SIG_N + 1];
} ko_qmodel_t;

static float row_scale(float maxabs) { return maxabs > 0.0f ? maxabs / 127.0f : 1.0f; }
static int8_t quant_w(float w, float s) {
    float q = rintf(w / s);
    if (q > 127.0f) q = 127.0f;
    if (q < -127.0f) q = -127.0f;
    return (int8_t) q;
}
static int32_t quant_mult(float s, int q_in) {
    double m = (double) s * ldexp(1.0, QP - q_in + 31);
    long long v = llrint(m);
    if (v > 2147483647LL) v = 2147483647LL;
    return (int32_t) v;
}
static int32_t quant_bias(float b) { return (int32_t) lrintf(b * 4096.0f); }
static float maxabs_col(const float *wt, int K, int N, int n, float c) {   /* max_k |c * W[n][k]| from the transposed copy Wt[k][n] */
    float m = 0.0f;
    for (int k = 0; k < K; k++) {
        float v = fabsf(c * wt[(size_t) k * N + n]);
        if (v > m) m = v;
    }
    return m;
}

static void qmodel_free(ko_qmodel_t *q) {
    if (!q) return;
    free(q->enc_q); free(q->enc_m); free(q->enc_b); free(q->dec_q); free(q->dec_m); free(q->dec_b);
    for (int l = 0; l < MAXL; l++) { free(q->ih_q[l]); free(q->hh_q[l]); free(q->g_m[l]); free(q->g_b[l]); }
    free(q);
}

static ko_qmodel_t *qmodel_build(const ko_model_t *m) {
    const int H = m->hidden;
    ko_qmodel_t *q = (ko_qmodel_t *) calloc(1, sizeof(*q));
    q->enc_q = (int8_t *) malloc((size_t) H * BINS);
    q->enc_m = (int32_t *) malloc(sizeof(int32_t) * H);
    q->enc_b = (int32_t *) malloc(sizeof(int32_t) * H);
    for (int n = 0; n < H; n++) {
        float s = row_scale(maxabs_col(m->enc_wt, BINS, H, n, 1.0f));
        for (int k = 0; k < BINS; k++) q->enc_q[(size_t) n * BINS + k] = quant_w(m->enc_wt[(size_t) k * H + n], s);
        q->enc_m[n] = quant_mult(s, QF);
        q->enc_b[n] = quant_bias(m->enc_b[n]);
    }
    for (int l = 0; l < m->layers; l++) {
        const float c = l == 0 ? 8.0f : 1.0f;   /* 2^(QH - Q of the layer input): e is Q12, h is Q15 */
        const float *wi = m->wih_t[l], *wh = m->whh_t[l];
        q->ih_q[l] = (int8_t *) malloc((size_t) 3 * H * H);
        q->hh_q[l] = (int8_t *) malloc((size_t) 3 * H * H);
        q->g_m[l] = (int32_t *) malloc(sizeof(int32_t) * 4 * H);
        q->g_b[l] = (int32_t *) malloc(sizeof(int32_t) * 4 * H);
        for (int j = 0; j < H; j++) {
            for (int g = 0; g < 2; g++) {       /* r, z: one scale for the x and the h half of the sum */
                const int n = g * H + j;
                float a = maxabs_col(wi, H, 3 * H, n, c), b = maxabs_col(wh, H, 3 * H, n, 1.0f);
                float s = row_scale(a > b ? a : b);
                for (int k = 0; k < H; k++) {
                    q->ih_q[l][(size_t) n * H + k] = quant_w(c * wi[(size_t) k * 3 * H + n], s);
                    q->hh_q[l][(size_t) n * H + k] = quant_w(wh[(size_t) k * 3 * H + n], s);
                }
                q->g_m[l][g * H + j] = quant_mult(s, QH);
                q->g_b[l][g * H + j] = quant_bias(m->bih[l][n] + m->bhh[l][n]);
            }
            const int n = 2 * H + j;
            float sx = row_scale(maxabs_col(wi, H, 3 * H, n, c)), sh = row_scale(maxabs_col(wh, H, 3 * H, n, 1.0f));
            for (int k = 0; k < H; k++) {
                q->ih_q[l][(size_t) n * H + k] = quant_w(c * wi[(size_t) k * 3 * H + n], sx);
                q->hh_q[l][(size_t) n * H + k] = quant_w(wh[(size_t) k * 3 * H + n], sh);
            }
            q->g_m[l][2 * H + j] = quant_mult(sx, QH);
            q->g_b[l][2 * H + j] = quant_bias(m->bih[l][n]);
            q->g_m[l][3 * H + j] = quant_mult(sh, QH);
            q->g_b[l][3 * H + j] = quant_bias(m->bhh[l][n]);
        }
    }
    q->dec_q = (int8_t *) malloc((size_t) BINS * H);
    q->dec_m = (int32_t *) malloc(sizeof(int32_t) * BINS);
    q->dec_b = (int32_t *) malloc(sizeof(int32_t) * BINS);
    for (int n = 0; n < BINS; n++) {
        float s = row_scale(maxabs_col(m->dec_wt, H, BINS, n, 1.0f));
        for (int k = 0; k < H; k++) q->dec_q[(size_t) n * H + k] = quant_w(m->dec_wt[(size_t) k * BINS + n], s);
        q->dec_m[n] = quant_mult(s, QH);
        q->dec_b[n] = quant_bias(m->dec_b[n]);
    }
    for (int i = 0; i <= SIG_N; i++) q->sig[i] = (int16_t) lrint(32768.0 / (1.0 + exp(-(double) (i - SIG_N / 2) / 128.0)));
    return q;
}

static inline int32_t requant(int32_t acc, int32_t mult) { return (int32_t) (((int64_t) acc * mult + ((int64_t) 1 << 30)) >> 31); }
static inline int32_t sig_q15(const int16_t *t, int32_t p) {        /* Q12 in, Q15 out */
    int32_t x = (p < -32768 ? -32768 : p > 32767 ? 32767 : p) + 32768;
    int32_t i = x >> 5, f = x & 31;
    return t[i] + ((((int32_t) t[i + 1] - t[i]) * f + 16) >> 5);
}
static inline int32_t tanh_q15(const int16_t *t, int32_t a) { return 2 * sig_q15(t, a < -16384 ? -32768 : a > 16383 ? 32767 : 2 * a) - 32768; }
static inline int32_t dot_q(const int8_t *w, const int16_t *x, int K) {   /* exact sum, reduced mod 2^32 like the int32 accumulator */
    int64_t s = 0;
    for (int k = 0; k < K; k++) s += (int32_t) w[k] * (int32_t) x[k];
    return (int32_t) (uint32_t) s;
}

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
    qmodel_free(m->q);
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
    m->q = qmodel_build(m);
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

/* fixed-point mask network of one stream (SPEC.md section 6): fq[256] int16 Q14 -> mask; the state is kept as h = q / 32768 */
KO_API void ko_masknet_q(ko_stream_t *s, const int16_t *fq, float *mask) {
    const ko_model_t *m = s->m;
    const ko_qmodel_t *q = m->q;
    const int H = m->hidden;
    int16_t *x = (int16_t *) malloc(sizeof(int16_t) * 2 * H), *hq = x + H;
    for (int n = 0; n < H; n++) {
        int32_t p = requant(dot_q(q->enc_q + (size_t) n * BINS, fq, BINS), q->enc_m[n]) + q->enc_b[n];
        x[n] = (int16_t) (p < 0 ? 0 : p > 32767 ? 32767 : p);
    }
    for (int l = 0; l < m->layers; l++) {
        float *h = s->h + l * H;
        const int32_t *gm = q->g_m[l], *gb = q->g_b[l];
        for (int j = 0; j < H; j++) hq[j] = (int16_t) lrintf(h[j] * 32768.0f);
        for (int j = 0; j < H; j++) {
            const int8_t *wi = q->ih_q[l], *wh = q->hh_q[l];
            int32_t ar = (int32_t) ((uint32_t) dot_q(wi + (size_t) j * H, x, H) + (uint32_t) dot_q(wh + (size_t) j * H, hq, H));
            int32_t az = (int32_t) ((uint32_t) dot_q(wi + (size_t) (H + j) * H, x, H) + (uint32_t) dot_q(wh + (size_t) (H + j) * H, hq, H));
            int32_t r = sig_q15(q->sig, requant(ar, gm[j]) + gb[j]);
            int32_t z = sig_q15(q->sig, requant(az, gm[H + j]) + gb[H + j]);
            int32_t pnx = requant(dot_q(wi + (size_t) (2 * H + j) * H, x, H), gm[2 * H + j]) + gb[2 * H + j];
            int32_t pnh = requant(dot_q(wh + (size_t) (2 * H + j) * H, hq, H), gm[3 * H + j]) + gb[3 * H + j];
            int32_t a = pnx + (int32_t) (((int64_t) r * pnh + (1 << 14)) >> 15);
            int32_t nn = tanh_q15(q->sig, a);
            int32_t hn = nn + (int32_t) (((int64_t) z * (hq[j] - nn) + (1 << 14)) >> 15);
            hn = hn < -32767 ? -32767 : hn > 32767 ? 32767 : hn;
            h[j] = (float) hn * (1.0f / 32768.0f);
        }
        for (int j = 0; j < H; j++) x[j] = (int16_t) lrintf(h[j] * 32768.0f);
    }
    for (int n = 0; n < BINS; n++) {
        int32_t p = requant(dot_q(q->dec_q + (size_t) n * H, x, H), q->dec_m[n]) + q->dec_b[n];
        mask[n] = (float) sig_q15(q->sig, p) * (1.0f / 32768.0f);
    }
    free(x);
    memcpy(s->last_mask, mask, sizeof(float) * BINS);
}

/* features -> int16 Q14, round to nearest even, saturating */
KO_API void ko_quantize_feat(const float *feat, int16_t *fq) {
    for (int k = 0; k < BINS; k++) {
        long v = lrintf(feat[k] * 16384.0f);
        fq[k] = (int16_t) (v < -32768 ? -32768 : v > 32767 ? 32767 : v);
    }
}

KO_API void ko_masknet(ko_stream_t *s, const float *feat, float *mask) {
    if (s->mode == 2) {
        int16_t fq[BINS];
        ko_quantize_feat(feat, fq);
        ko_masknet_q(s, fq, mask);
        return;
    }
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
            if (j->b->mode == 2) for (int s = 0; s < S; s++) ko_masknet(j->b->st[s0 + s], feat + s * BINS, mask + s * BINS);
            else masknet_block(j->b->st + s0, S, feat, mask, scratch);
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
