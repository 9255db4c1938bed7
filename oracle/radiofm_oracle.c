/* oracle/radiofm_oracle.c -- plain-C restatement of the reference IQ->audio(+RDS) DSP chain.
 *
 * TEST INFRASTRUCTURE ONLY (see radiofm_oracle.h).  Independent of the product sources.
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/src).  Precision rules follow SURVEY.md Appendix C: RealType is float32;
 * a sub-expression is evaluated in double wherever a double literal or K_* macro takes part
 * and is rounded to float32 when stored; no FMA contraction (build with -ffp-contract=off);
 * x87 fsincos on a float phase == float(sin/cos(double phase)).
 *
 * Parity pin: bit-exact against oracle/_ref/libradiofm_ref.so (the unmodified reference) in
 * tests/test_oracle_port.py, on every stage tap, for all BASELINE.json rates.
 */
#include "radiofm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define K_2PI (2.0 * 3.14159265358979323846)
#define K_PI (3.14159265358979323846)
#define K_PI2 (K_PI / 2.0)

typedef struct { float re, im; } cf32;

/* ------------------------------------------------------------------------------------------
 * libm restatements
 * ---------------------------------------------------------------------------------------- */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* glibc 2.39 sysdeps/ieee754/flt-32/s_atanf.c (fdlibm, float arithmetic, no FMA).  Pinned
 * against libm atan2f over 2e8 random arguments + the full float range of atanf in
 * tests/test_oracle_port.py. */
static const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
static const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
static const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                             9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                             4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};

static float rfo_atanf(float x)
{
  int32_t hx = (int32_t)f2u(x), ix = hx & 0x7fffffff;
  int id;
  float z, w, s1, s2;
  if (ix >= 0x4c000000) { /* |x| >= 2^25 */
    if (ix > 0x7f800000) return x + x;
    if (hx > 0) return atanhi[3] + atanlo[3];
    return -atanhi[3] - atanlo[3];
  }
  if (ix < 0x3ee00000) { /* |x| < 0.4375 */
    if (ix < 0x31000000) return x; /* |x| < 2^-29 */
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) { /* |x| < 1.1875 */
      if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
      else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
      else { id = 3; x = -1.0f / x; }
    }
  }
  z = x * x;
  w = z * z;
  s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
  return (hx < 0) ? -z : z;
}

/* glibc 2.39 sysdeps/ieee754/flt-32/e_atan2f.c */
float rfo_atan2f(float y, float x)
{
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f,
              pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  int32_t hx = (int32_t)f2u(x), ix = hx & 0x7fffffff, hy = (int32_t)f2u(y), iy = hy & 0x7fffffff;
  int m, k;
  float z;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return rfo_atanf(y);
  m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) { case 0: case 1: return y; case 2: return pi + tiny; default: return -pi - tiny; }
  }
  if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) { case 0: return pi_o_4 + tiny; case 1: return -pi_o_4 - tiny;
                   case 2: return 3.0f * pi_o_4 + tiny; default: return -3.0f * pi_o_4 - tiny; }
    } else {
      switch (m) { case 0: return 0.0f; case 1: return -0.0f; case 2: return pi + tiny; default: return -pi - tiny; }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  k = (iy - ix) >> 23;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = rfo_atanf(fabsf(y / x));
  switch (m) {
    case 0: return z;
    case 1: return u2f(f2u(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

/* x87 `fsincos` on a float32 phase with the result rounded to float32 (FmDecode.cpp:167,386,
 * RDSProcess.cpp:245, FreqShift.cpp:56): identical to float(sin/cos(double)) in 1e8 random
 * phases. */
void rfo_sincos(float phase, float* s, float* c)
{
  *s = (float)sin((double)phase);
  *c = (float)cos((double)phase);
}

/* ------------------------------------------------------------------------------------------
 * cFineTuner -- FmDecode.cpp:45-82
 * ---------------------------------------------------------------------------------------- */
typedef struct { unsigned index, size; cf32* table; } finetuner;

static void finetuner_init(finetuner* t, unsigned table_size, int freq_shift)
{
  t->index = 0;
  t->size = table_size;
  t->table = (cf32*)calloc(table_size, sizeof(cf32));
  float phase_step = (float)(K_2PI / (float)table_size);           /* :50 */
  for (unsigned i = 0; i < table_size; ++i) {
    int64_t r = ((int64_t)freq_shift * (int64_t)i) % (int64_t)table_size; /* :53 sign of dividend */
    float phi = (float)r * phase_step;
    float pcos = cosf(phi), psin = sinf(phi);
    t->table[i].re = pcos * 2.0f;                                   /* :56 */
    t->table[i].im = psin * 2.0f;
  }
}

static void finetuner_process(finetuner* t, const cf32* in, cf32* out, unsigned n)
{
  unsigned idx = t->index;
  for (unsigned i = 0; i < n; ++i) {                                /* :71-77 complex product */
    cf32 a = in[i], b = t->table[idx];
    out[i].re = a.re * b.re - a.im * b.im;
    out[i].im = a.re * b.im + a.im * b.re;
    if (++idx == t->size) idx = 0;
  }
  t->index = idx;
}

/* ------------------------------------------------------------------------------------------
 * cDownsampleFilter -- DownConvert.cpp:18-256
 * ---------------------------------------------------------------------------------------- */
struct rfo_downsample {
  double downsample;
  unsigned downsample_int, pos_int, order;
  float pos_frac;
  float* coeff;     /* order + 2 entries, [0] = [order+1] = 0 */
  float* state_r;   /* order */
  cf32* state_c;    /* order */
};

static float* make_lanczos(unsigned filter_order, double cutoff)   /* :18-56 */
{
  float* coeff = (float*)calloc(filter_order + 3, sizeof(float));
  double ysum = 0.0;
  for (int i = 1; i <= (int)filter_order + 1; i++) {
    int t2 = 2 * i - (int)filter_order;
    double y;
    if (t2 == 0) {
      y = 1.0;
    } else {
      double x1 = cutoff * t2;
      double x2 = t2 / (double)(filter_order + 2);
      y = ((double)sinf((float)(K_PI * x1)) / K_PI / x1) * ((double)sinf((float)(K_PI * x2)) / K_PI / x2);
    }
    coeff[i] = (float)y;
    ysum += y;
  }
  for (unsigned i = 1; i <= filter_order + 1; i++)
    coeff[i] = (float)((double)coeff[i] / ysum);
  return coeff;
}

static void downsample_init(rfo_downsample* f, unsigned order, double cutoff, double downsample, int integer_factor)
{
  f->downsample = downsample;
  f->downsample_int = integer_factor ? (unsigned)lrint(downsample) : 0;
  f->pos_int = 0;
  f->pos_frac = 0;
  f->order = order;
  f->coeff = make_lanczos(order - 1, cutoff);                       /* :78 */
  f->state_c = (cf32*)calloc(order ? order : 1, sizeof(cf32));
  f->state_r = (float*)calloc(order ? order : 1, sizeof(float));
}

static void downsample_free(rfo_downsample* f)
{
  free(f->coeff); free(f->state_c); free(f->state_r);
}

/* complex, integer factor: :98-154.  Written over the virtual sequence V = state ++ input. */
static unsigned downsample_complex(rfo_downsample* f, const cf32* in, cf32* out, unsigned n)
{
  const unsigned order = f->order, pstep = f->downsample_int;
  unsigned p = f->pos_int, i = 0;
  for (; p < n; p += pstep, i++) {
    float yr = 0, yi = 0;
    for (unsigned j = 1; j <= order; j++) {
      cf32 s = (j <= p) ? in[p - j] : f->state_c[order + p - j];
      yr += s.re * f->coeff[j];
      yi += s.im * f->coeff[j];
    }
    out[i].re = yr; out[i].im = yi;
  }
  f->pos_int = p - n;
  if (n < order) {
    memmove(f->state_c, f->state_c + n, (order - n) * sizeof(cf32));
    memcpy(f->state_c + (order - n), in, n * sizeof(cf32));
  } else {
    memcpy(f->state_c, in + (n - order), order * sizeof(cf32));
  }
  return i;
}

/* real: integer branch :166-194, fractional branch :195-233 */
static unsigned downsample_real(rfo_downsample* f, const float* in, float* out, unsigned n)
{
  const unsigned order = f->order;
  unsigned i = 0;
  if (f->downsample_int != 0) {
    const unsigned pstep = f->downsample_int;
    unsigned p = f->pos_int;
    for (; p < n; p += pstep, i++) {
      float y = 0;
      for (unsigned j = 1; j <= order; j++) {
        float s = (j <= p) ? in[p - j] : f->state_r[order + p - j];
        y += s * f->coeff[j];
      }
      out[i] = y;
    }
    f->pos_int = p - n;
  } else {
    float p = f->pos_frac;
    float pstep = (float)f->downsample;
    float pf = p;
    unsigned pi = (unsigned)(int)pf;
    while (pi < n) {
      float k1 = pf - (float)pi;
      float k0 = 1 - k1;
      float y = 0;
      for (unsigned j = 0; j <= order; j++) {
        float k = f->coeff[j] * k0 + f->coeff[j + 1] * k1;
        float s = (j <= pi) ? in[pi - j] : f->state_r[order + pi - j];
        y += k * s;
      }
      out[i] = y;
      i++;
      pf = p + (float)i * pstep;
      pi = (unsigned)(int)pf;
    }
    f->pos_frac = pf - (float)n;
    if (f->pos_frac < 0) f->pos_frac = 0;
  }
  if (n < order) {
    memmove(f->state_r, f->state_r + n, (order - n) * sizeof(float));
    memcpy(f->state_r + (order - n), in, n * sizeof(float));
  } else {
    memcpy(f->state_r, in + (n - order), order * sizeof(float));
  }
  return i;
}

/* ------------------------------------------------------------------------------------------
 * cFirFilter -- FirFilter.cpp
 * ---------------------------------------------------------------------------------------- */
#define MAX_NUMCOEF 75
struct rfo_fir {
  float fs;
  unsigned ntaps;
  int state;
  float coef[MAX_NUMCOEF], icoef[MAX_NUMCOEF], qcoef[MAX_NUMCOEF];
  float rz[MAX_NUMCOEF];
  cf32 cz[MAX_NUMCOEF];
};

static float izero(float x)                                         /* :39-58 */
{
  float x2 = x / 2.0f, sum = 1.0f, ds = 1.0f, di = 1.0f, errorlimit = (float)1e-9, tmp;
  do {
    tmp = x2 / di;
    tmp *= tmp;
    ds *= tmp;
    sum += ds;
    di = (float)(di + 1.0);
  } while (ds >= errorlimit * sum);
  return sum;
}

static void fir_clear(rfo_fir* f)
{
  memset(f->rz, 0, sizeof(f->rz));
  memset(f->cz, 0, sizeof(f->cz));
  f->state = 0;
}

static void fir_ctor(rfo_fir* f) { memset(f, 0, sizeof(*f)); f->ntaps = 1; }

static int fir_init_lp(rfo_fir* f, unsigned NumTaps, float Scale, float Astop, float Fpass, float Fstop, float Fs)
{                                                                   /* :78-148 */
  float Beta;
  f->fs = Fs;
  float normFpass = Fpass / Fs, normFstop = Fstop / Fs;
  float normFcut = (normFstop + normFpass) / 2.0f;
  if (Astop < 20.96f) Beta = 0;
  else if (Astop >= 50.0f) Beta = (float)(.1102 * (Astop - 8.71f));
  else Beta = (float)(.5842 * powf((Astop - 20.96f), (float)0.4) + .07886f * (Astop - 20.96f));
  f->ntaps = (unsigned)((Astop - 8.0f) / (2.285f * K_2PI * (normFstop - normFpass)) + 1);
  if (f->ntaps > MAX_NUMCOEF) f->ntaps = MAX_NUMCOEF;
  if (f->ntaps < 3) f->ntaps = 3;
  if (NumTaps) f->ntaps = NumTaps;
  float fCenter = (float)(.5 * (float)(f->ntaps - 1));
  float izb = izero(Beta);
  for (unsigned n = 0; n < f->ntaps; ++n) {
    float x = (float)n - fCenter;
    float c;
    if ((float)n == fCenter) c = (float)(2.0 * normFcut);
    else c = (float)(sinf((float)(K_2PI * x * normFcut)) / (K_PI * x));
    x = ((float)n - ((float)f->ntaps - 1.0f) / 2.0f) / (((float)f->ntaps - 1.0f) / 2.0f);
    f->coef[n] = Scale * c * izero(Beta * sqrtf(1 - (x * x))) / izb;
  }
  for (unsigned n = 0; n < f->ntaps; ++n) { f->icoef[n] = f->coef[n]; f->qcoef[n] = f->coef[n]; }
  fir_clear(f);
  return (int)f->ntaps;
}

static int fir_init_hp(rfo_fir* f, unsigned NumTaps, float Scale, float Astop, float Fpass, float Fstop, float Fs)
{                                                                   /* :195-264 */
  float Beta;
  f->fs = Fs;
  float normFpass = Fpass / Fs, normFstop = Fstop / Fs;
  float normFcut = (float)((normFstop + normFpass) / 2.0);
  if (Astop < 20.96f) Beta = 0;
  else if (Astop >= 50.0f) Beta = .1102f * (Astop - 8.71f);
  else Beta = .5842f * powf((Astop - 20.96f), 0.4f) + .07886f * (Astop - 20.96f);
  f->ntaps = (unsigned)((Astop - 8.0f) / (2.285f * K_2PI * (normFpass - normFstop)) + 1);
  if (f->ntaps > (MAX_NUMCOEF - 1)) f->ntaps = MAX_NUMCOEF - 1;
  if (f->ntaps < 3) f->ntaps = 3;
  f->ntaps |= 1;
  if (NumTaps) f->ntaps = NumTaps;
  float izb = izero(Beta);
  float fCenter = .5f * (float)(f->ntaps - 1);
  for (unsigned n = 0; n < f->ntaps; n++) {
    float x = (float)((float)n - (float)(f->ntaps - 1) / 2.0);
    float c;
    if ((float)n == fCenter) c = (float)(1.0 - 2.0 * normFcut);
    else c = (float)(sinf((float)(K_PI * x)) / (K_PI * x) - sinf((float)(K_2PI * x * normFcut)) / (K_PI * x));
    x = ((float)n - ((float)f->ntaps - 1.0f) / 2.0f) / (((float)f->ntaps - 1.0f) / 2.0f);
    f->coef[n] = Scale * c * izero(Beta * sqrtf(1 - (x * x))) / izb;
  }
  for (unsigned n = 0; n < f->ntaps; ++n) { f->icoef[n] = f->coef[n]; f->qcoef[n] = f->coef[n]; }
  fir_clear(f);
  return (int)f->ntaps;
}

static void fir_init_const(rfo_fir* f, unsigned NumTaps, const float* coef, float Fs)
{                                                                   /* :302-320 (m_Coef only) */
  f->fs = Fs;
  f->ntaps = NumTaps > MAX_NUMCOEF ? MAX_NUMCOEF : NumTaps;
  for (unsigned i = 0; i < f->ntaps; ++i) f->coef[i] = coef[i];
  fir_clear(f);
}

/* Circular delay line Z[state] = newest; tap index for Z[j] is (j - state) mod N; the sum
 * runs j = 0..N-1, i.e. it STARTS at a tap that rotates with m_State (:330-413). */
static void fir_process_real(rfo_fir* f, float* buf, unsigned n)    /* :360-377 */
{
  const int N = (int)f->ntaps;
  for (unsigned i = 0; i < n; ++i) {
    f->rz[f->state] = buf[i];
    int h = (N - f->state) % N;
    float acc = f->coef[h] * f->rz[0];
    for (int j = 1; j < N; ++j) { if (++h == N) h = 0; acc += f->coef[h] * f->rz[j]; }
    if (--f->state < 0) f->state += N;
    buf[i] = acc;
  }
}

static void fir_process_complex(rfo_fir* f, cf32* buf, unsigned n)  /* :330-350 */
{
  const int N = (int)f->ntaps;
  for (unsigned i = 0; i < n; ++i) {
    f->cz[f->state] = buf[i];
    int h = (N - f->state) % N;
    float ar = f->icoef[h] * f->cz[0].re, ai = f->qcoef[h] * f->cz[0].im;
    for (int j = 1; j < N; ++j) {
      if (++h == N) h = 0;
      ar += f->icoef[h] * f->cz[j].re;
      ai += f->qcoef[h] * f->cz[j].im;
    }
    if (--f->state < 0) f->state += N;
    buf[i].re = ar; buf[i].im = ai;
  }
}

static void fir_process_two(rfo_fir* f, float* a, float* b, unsigned n) /* :387-413 */
{
  const int N = (int)f->ntaps;
  for (unsigned i = 0; i < n; ++i) {
    f->cz[f->state].re = a[i]; f->cz[f->state].im = b[i];
    int h = (N - f->state) % N;
    float va = f->icoef[h] * f->cz[0].re, vb = f->qcoef[h] * f->cz[0].im;
    for (int j = 1; j < N; ++j) {
      if (++h == N) h = 0;
      va += f->icoef[h] * f->cz[j].re;
      vb += f->qcoef[h] * f->cz[j].im;
    }
    if (--f->state < 0) f->state += N;
    a[i] = va; b[i] = vb;
  }
}

/* ------------------------------------------------------------------------------------------
 * cIirFilter -- IirFilter.cpp
 * ---------------------------------------------------------------------------------------- */
struct rfo_iir { float A1, A2, B0, B1, B2, w1a, w2a, w1b, w2b; };

static int iir_init(rfo_iir* f, int type, float F0, float Q, float Fs)  /* :11-60 */
{
  int ret = 1;
  float w0 = (float)(K_2PI * F0 / Fs);
  float alpha = (float)(sinf(w0) / (2.0 * Q));
  float A = (float)(1.0 / (1.0 + alpha));
  switch (type) {
    case 0: /* ftLP */
      f->B0 = (float)(A * ((1.0 - cosf(w0)) / 2.0));
      f->B1 = (float)(A * (1.0 - cosf(w0)));
      f->B2 = (float)(A * ((1.0 - cosf(w0)) / 2.0));
      f->A1 = (float)(A * (-2.0 * cosf(w0)));
      f->A2 = (float)(A * (1.0 - alpha));
      break;
    case 1: /* ftHP */
      f->B0 = (float)(A * ((1.0 + cosf(w0)) / 2.0));
      f->B1 = (float)(-A * (1.0 + cosf(w0)));
      f->B2 = (float)(A * ((1.0 + cosf(w0)) / 2.0));
      f->A1 = (float)(A * (-2.0 * cosf(w0)));
      f->A2 = (float)(A * (1.0 - alpha));
      break;
    case 2: /* ftBP */
      f->B0 = A * alpha;
      f->B1 = 0.0f;
      f->B2 = A * -alpha;
      f->A1 = (float)(A * (-2.0 * cosf(w0)));
      f->A2 = (float)(A * (1.0 - alpha));
      break;
    case 3: /* ftBR */
      f->B0 = (float)(A * 1.0);
      f->B1 = (float)(A * (-2.0 * cosf(w0)));
      f->B2 = (float)(A * 1.0);
      f->A1 = (float)(A * (-2.0 * cosf(w0)));
      f->A2 = (float)(A * (1.0 - alpha));
      break;
    default:
      ret = 0;
      break;
  }
  f->w1a = f->w2a = f->w1b = f->w2b = 0.0f;
  return ret;
}

static inline float biquad_step(const rfo_iir* f, float x, float* w1, float* w2) /* :78-105 */
{
  float w0 = x - f->A1 * *w1 - f->A2 * *w2;
  float y = f->B0 * w0 + f->B1 * *w1 + f->B2 * *w2;
  *w2 = *w1;
  *w1 = w0;
  return y;
}

static void iir_process_real(rfo_iir* f, float* buf, unsigned n)
{
  for (unsigned i = 0; i < n; ++i) buf[i] = biquad_step(f, buf[i], &f->w1a, &f->w2a);
}

static void iir_process_two(rfo_iir* f, float* a, float* b, unsigned n)
{
  for (unsigned i = 0; i < n; ++i) {
    a[i] = biquad_step(f, a[i], &f->w1a, &f->w2a);
    b[i] = biquad_step(f, b[i], &f->w1b, &f->w2b);
  }
}

/* ------------------------------------------------------------------------------------------
 * cPilotPhaseLock -- FmDecode.cpp:88-229
 * ---------------------------------------------------------------------------------------- */
struct rfo_pilot {
  float minfreq, maxfreq, b0, a1, a2, i1, i2, q1, q2, lb0, lb1, x1, freq, phase, minsignal, level;
  int lock_delay, lock_cnt;
};

static void pilot_init(rfo_pilot* p, float freq, float bandwidth, float minsignal)
{
  p->minfreq = (float)((freq - bandwidth) * K_2PI);
  p->maxfreq = (float)((freq + bandwidth) * K_2PI);
  p->minsignal = minsignal;
  p->lock_delay = (int)(20.0f / bandwidth);
  p->lock_cnt = 0;
  float p1 = (float)exp(-1.146f * bandwidth * K_2PI);
  float p2 = (float)exp(-5.331f * bandwidth * K_2PI);
  p->a1 = -p1 - p2;
  p->a2 = p1 * p2;
  p->b0 = 1 + p->a1 + p->a2;
  p->lb0 = (float)(0.62f * bandwidth * K_2PI);
  p->lb1 = (float)(-p->lb0 * exp(-0.1153 * bandwidth * K_2PI));
  p->freq = (float)(freq * K_2PI);
  p->phase = 0;
  p->i1 = p->i2 = p->q1 = p->q2 = p->x1 = 0;
  p->level = 0;
}

static int pilot_process(rfo_pilot* p, const float* in, float* out, unsigned n)
{
  p->level = 1000.0f;
  for (unsigned i = 0; i < n; ++i) {
    float psin, pcos;
    rfo_sincos(p->phase, &psin, &pcos);
    out[i] = 2 * psin * pcos;
    float x = in[i];
    float pi_ = psin * x, pq = pcos * x;
    pi_ = p->b0 * pi_ - p->a1 * p->i1 - p->a2 * p->i2;
    pq = p->b0 * pq - p->a1 * p->q1 - p->a2 * p->q2;
    p->i2 = p->i1; p->i1 = pi_;
    p->q2 = p->q1; p->q1 = pq;
    float err;
    if (pi_ > fabsf(pq)) err = pq / pi_;
    else if (pq > 0) err = 1;
    else err = -1;
    p->level = (pi_ < p->level) ? pi_ : p->level;                  /* std::min(level, phasor_i) */
    p->freq += p->lb0 * err + p->lb1 * p->x1;
    p->x1 = err;
    { float t = (p->freq < p->maxfreq) ? p->freq : p->maxfreq;     /* std::min(max, f) */
      p->freq = (p->minfreq < t) ? t : p->minfreq; }               /* std::max(min, t) */
    p->phase += p->freq;
    if (p->phase > K_2PI) p->phase = (float)(p->phase - K_2PI);
  }
  if (2 * p->level > p->minsignal) {
    if (p->lock_cnt < p->lock_delay) p->lock_cnt += (int)n;
  } else {
    p->lock_cnt = 0;
  }
  return p->lock_cnt >= p->lock_delay;
}

/* ------------------------------------------------------------------------------------------
 * CRDSDownConvert -- DownConvert.cpp:271-727, taps from filtercoef.h:62-150 (BSD, Moe Wheatley;
 * reused as data).  The tap tables below are the published half-band designs.
 * ---------------------------------------------------------------------------------------- */
#include "halfband_taps.inc"

#define MAX_DECSTAGES 10
typedef struct {
  int len;           /* 3 = CIC3, 11 = fixed 11-tap, else generic half-band length */
  const float* h;
  cf32* tail;        /* len-1 previous inputs (generic) / 10 (11-tap) / Xodd,Xeven (CIC3) */
} dec2_stage;

struct rfo_rdsdc {
  float out_rate, nco_freq, cw_offset, nco_inc, in_rate, max_bw;
  cf32 osc1;
  float osc_cos, osc_sin;
  int nstages;
  dec2_stage st[MAX_DECSTAGES];
};

static void rdsdc_delete(rfo_rdsdc* d)
{
  for (int i = 0; i < d->nstages; ++i) free(d->st[i].tail);
  d->nstages = 0;
}

static void rdsdc_add(rfo_rdsdc* d, int len, const float* h)
{
  dec2_stage* s = &d->st[d->nstages++];
  s->len = len; s->h = h;
  s->tail = (cf32*)calloc((size_t)(len - 1), sizeof(cf32));
}

static void rdsdc_init(rfo_rdsdc* d)                                /* :271-285 */
{
  memset(d, 0, sizeof(*d));
  d->in_rate = 100000.0f; d->max_bw = 10000.0f;
  d->osc1.re = 1.0f; d->osc1.im = 0.0f;
}

static void rdsdc_set_frequency(rfo_rdsdc* d, float f)              /* :311-320 */
{
  float tmpf = f + d->cw_offset;
  d->nco_freq = tmpf;
  d->nco_inc = (float)(K_2PI * d->nco_freq / d->in_rate);
  d->osc_cos = cosf(d->nco_inc);
  d->osc_sin = sinf(d->nco_inc);
}

static float rdsdc_set_data_rate(rfo_rdsdc* d, float InRate, float MaxBW) /* :327-371 */
{
  float f = InRate;
  if ((d->in_rate != InRate) || (d->max_bw != MaxBW)) {
    d->in_rate = InRate; d->max_bw = MaxBW;
    rdsdc_delete(d);
    while ((f > (d->max_bw / HB51TAP_MAX)) && (f > (7900.0 * 2.0))) {
      if (f >= (d->max_bw / CIC3_MAX)) rdsdc_add(d, 3, NULL);
      else if (f >= (d->max_bw / HB11TAP_MAX)) rdsdc_add(d, 11, NULL);
      else if (f >= (d->max_bw / HB15TAP_MAX)) rdsdc_add(d, 15, HB15TAP_H);
      else if (f >= (d->max_bw / HB19TAP_MAX)) rdsdc_add(d, 19, HB19TAP_H);
      else if (f >= (d->max_bw / HB23TAP_MAX)) rdsdc_add(d, 23, HB23TAP_H);
      else if (f >= (d->max_bw / HB27TAP_MAX)) rdsdc_add(d, 27, HB27TAP_H);
      else if (f >= (d->max_bw / HB31TAP_MAX)) rdsdc_add(d, 31, HB31TAP_H);
      else if (f >= (d->max_bw / HB35TAP_MAX)) rdsdc_add(d, 35, HB35TAP_H);
      else if (f >= (d->max_bw / HB39TAP_MAX)) rdsdc_add(d, 39, HB39TAP_H);
      else if (f >= (d->max_bw / HB43TAP_MAX)) rdsdc_add(d, 43, HB43TAP_H);
      else if (f >= (d->max_bw / HB47TAP_MAX)) rdsdc_add(d, 47, HB47TAP_H);
      else if (f >= (d->max_bw / HB51TAP_MAX)) rdsdc_add(d, 51, HB51TAP_H);
      f = (float)(f / 2.0);
    }
    d->out_rate = f;
    rdsdc_set_frequency(d, d->nco_freq);
  }
  return d->out_rate;
}

static float rdsdc_set_wfm_data_rate(rfo_rdsdc* d, float InRate, float MaxBW) /* :378-399 */
{
  float f = InRate;
  if ((d->in_rate != InRate) || (d->max_bw != MaxBW)) {
    d->in_rate = InRate; d->max_bw = MaxBW;
    rdsdc_delete(d);
    while (f > 400000.0) { rdsdc_add(d, 51, HB51TAP_H); f = (float)(f / 2.0); }
    d->out_rate = f;
    rdsdc_set_frequency(d, d->nco_freq);
  }
  return d->out_rate;
}

#define MAX_HALF_BAND_BUFSIZE 32768
/* generic half-band: :516-550 (tap 0 applied twice: acc = b[i]*H[0]; then j = 0,2,...) */
static int dec2_halfband(dec2_stage* s, int n, cf32* io)
{
  const int L = s->len;
  if (n < L) return n / 2;                                          /* :519-520 "safety net" */
  cf32* buf = (cf32*)malloc((size_t)(n + L) * sizeof(cf32));
  memcpy(buf, s->tail, (size_t)(L - 1) * sizeof(cf32));
  memcpy(buf + (L - 1), io, (size_t)n * sizeof(cf32));
  int k = 0;
  for (int i = 0; i < n; i += 2) {
    float ar = buf[i].re * s->h[0], ai = buf[i].im * s->h[0];
    for (int j = 0; j < L; j += 2) {
      ar = ar + buf[i + j].re * s->h[j];
      ai = ai + buf[i + j].im * s->h[j];
    }
    ar = ar + buf[i + (L - 1) / 2].re * s->h[(L - 1) / 2];
    ai = ai + buf[i + (L - 1) / 2].im * s->h[(L - 1) / 2];
    io[k].re = ar; io[k].im = ai; k++;
  }
  memcpy(s->tail, buf + n, (size_t)(L - 1) * sizeof(cf32));
  free(buf);
  return k;
}

/* fixed 11-tap: :589-688 -- out[k] = H0 V[2k] + H2 V[2k+2] + H4 V[2k+4] + H5 V[2k+5] + H6 V[2k+6]
 * + H8 V[2k+8] + H10 V[2k+10] summed left to right over V = d0..d9 ++ input */
static int dec2_hb11(dec2_stage* s, int n, cf32* io)
{
  const float H0 = HB11TAP_H[0], H2 = HB11TAP_H[2], H4 = HB11TAP_H[4], H5 = HB11TAP_H[5],
              H6 = HB11TAP_H[6], H8 = HB11TAP_H[8], H10 = HB11TAP_H[10];
  cf32* v = (cf32*)malloc((size_t)(n + 10) * sizeof(cf32));
  memcpy(v, s->tail, 10 * sizeof(cf32));
  memcpy(v + 10, io, (size_t)n * sizeof(cf32));
  for (int k = 0; k < n / 2; ++k) {
    const cf32* w = v + 2 * k;
    io[k].re = H0 * w[0].re + H2 * w[2].re + H4 * w[4].re + H5 * w[5].re + H6 * w[6].re + H8 * w[8].re + H10 * w[10].re;
    io[k].im = H0 * w[0].im + H2 * w[2].im + H4 * w[4].im + H5 * w[5].im + H6 * w[6].im + H8 * w[8].im + H10 * w[10].im;
  }
  memcpy(s->tail, v + n, 10 * sizeof(cf32));
  free(v);
  return n / 2;
}

/* CIC N=3: :709-727; tail[0] = Xodd, tail[1] = Xeven */
static int dec2_cic3(dec2_stage* s, int n, cf32* io)
{
  int j = 0;
  for (int i = 0; i < n; i += 2, j++) {
    cf32 even = io[i], odd = io[i + 1];
    io[j].re = (float)(.125 * ((odd.re + s->tail[1].re) + 3.0 * (s->tail[0].re + even.re)));
    io[j].im = (float)(.125 * ((odd.im + s->tail[1].im) + 3.0 * (s->tail[0].im + even.im)));
    s->tail[0] = odd; s->tail[1] = even;
  }
  return j;
}

static int rdsdc_process(rfo_rdsdc* d, int n, cf32* in, cf32* out)  /* :412-489, NCO_OSC */
{
  for (int i = 0; i < n; i++) {
    cf32 dtmp = in[i], osc;
    osc.re = d->osc1.re * d->osc_cos - d->osc1.im * d->osc_sin;
    osc.im = d->osc1.im * d->osc_cos + d->osc1.re * d->osc_sin;
    float gn = (float)(1.95 - (d->osc1.re * d->osc1.re + d->osc1.im * d->osc1.im));
    d->osc1.re = gn * osc.re;
    d->osc1.im = gn * osc.im;
    in[i].re = (dtmp.re * osc.re) - (dtmp.im * osc.im);
    in[i].im = (dtmp.re * osc.im) + (dtmp.im * osc.re);
  }
  int m = n;
  for (int j = 0; j < d->nstages; ++j) {
    dec2_stage* s = &d->st[j];
    if (s->len == 3) m = dec2_cic3(s, m, in);
    else if (s->h == NULL) m = dec2_hb11(s, m, in);
    else m = dec2_halfband(s, m, in);
  }
  for (int i = 0; i < m; i++) out[i] = in[i];
  return m;
}

/* ------------------------------------------------------------------------------------------
 * RDS block sync + FEC -- RDSProcess.cpp:13-41,272-431 (integer)
 * ---------------------------------------------------------------------------------------- */
static const int BLK_OFFSET_TBL[8] = {0x3D8, 0x3D4, 0x25C, 0x258, 0x3D8, 0x3D4, 0x3CC, 0x258};
static const uint32_t PARCKH[16] = {0x2DC, 0x16E, 0x0B7, 0x287, 0x39F, 0x313, 0x355, 0x376,
                                    0x1BB, 0x201, 0x3DC, 0x1EE, 0x0F7, 0x2A7, 0x38F, 0x31B};
enum { ST_BITSYNC = 0, ST_BLOCKSYNC = 1, ST_GROUPDECODE = 2, ST_GROUPRESYNC = 3 };

struct rfo_rdssync {
  uint32_t in_bits;
  int cur_block, bit_pos, state, bgroup_off, block_errors;
  uint16_t block[4];
  uint16_t* groups; unsigned ngroups, cap;
};

static uint32_t check_block(uint32_t* in_bits, uint32_t SyndromeOffset, int UseFec) /* :377-431 */
{
  uint32_t testblock = (0x3FFFFFF & *in_bits);
  uint32_t syndrome = testblock >> 16;
  for (int i = 0; i < 16; i++) {
    if (testblock & 0x8000) syndrome ^= PARCKH[i];
    testblock <<= 1;
  }
  syndrome ^= SyndromeOffset;
  if (syndrome && UseFec) {
    uint32_t correctmask = (1u << 25);
    for (int i = 0; i < 16; i++) {
      if (syndrome & 0x200) {
        if (0 == (syndrome & 0x1F)) { *in_bits ^= correctmask; syndrome <<= 1; }
        else { syndrome <<= 1; syndrome ^= 0x5B9; }
      } else {
        syndrome <<= 1;
      }
      correctmask >>= 1;
    }
    syndrome &= 0x3FF;
  }
  return syndrome;
}

uint32_t rfo_rds_check_block(uint32_t word26, uint32_t offset_syndrome, int use_fec, uint32_t* corrected)
{
  uint32_t w = word26;
  uint32_t s = check_block(&w, offset_syndrome, use_fec);
  if (corrected) *corrected = w;
  return s;
}

static void rdssync_emit(rfo_rdssync* s)
{
  if (s->ngroups == s->cap) {
    s->cap = s->cap ? 2 * s->cap : 64;
    s->groups = (uint16_t*)realloc(s->groups, (size_t)s->cap * 4 * sizeof(uint16_t));
  }
  memcpy(s->groups + 4 * (size_t)s->ngroups, s->block, 4 * sizeof(uint16_t));
  s->ngroups++;
}

static void rdssync_reset(rfo_rdssync* s)                           /* RDSProcess.cpp:108-118 */
{
  s->bit_pos = 0; s->cur_block = 0; s->state = ST_BITSYNC; s->bgroup_off = 0;
}

static void rdssync_bit(rfo_rdssync* s, int bit)                    /* :272-375 */
{
  s->in_bits = (s->in_bits << 1) | (uint32_t)bit;
  switch (s->state) {
    case ST_BITSYNC:
      if (!check_block(&s->in_bits, 0x3D8, 0)) {
        s->bit_pos = 0; s->bgroup_off = 0;
        s->block[0] = (uint16_t)(s->in_bits >> 10);
        s->cur_block = 1; s->state = ST_BLOCKSYNC;
      }
      break;
    case ST_BLOCKSYNC:
      if (++s->bit_pos >= 26) {
        s->bit_pos = 0;
        if (check_block(&s->in_bits, (uint32_t)BLK_OFFSET_TBL[s->cur_block + s->bgroup_off], 0)) {
          s->state = ST_BITSYNC;
        } else {
          s->block[s->cur_block] = (uint16_t)(s->in_bits >> 10);
          if (s->cur_block == 1 && (s->block[1] & 0x0800)) s->bgroup_off = 4; else s->bgroup_off = 0;
          if (s->cur_block >= 3) {
            s->cur_block = 0; s->block_errors = 0; s->state = ST_GROUPDECODE;
            rdssync_emit(s);
          } else {
            s->cur_block++;
          }
        }
      }
      break;
    case ST_GROUPDECODE:
      if (++s->bit_pos >= 26) {
        s->bit_pos = 0;
        if (check_block(&s->in_bits, (uint32_t)BLK_OFFSET_TBL[s->cur_block + s->bgroup_off], 1)) {
          s->block_errors++;
          if (s->block_errors > 0) {
            s->state = ST_BITSYNC;
          } else {
            if (++s->cur_block > 3) s->cur_block = 0;
            if (s->cur_block != 0) s->state = ST_GROUPRESYNC;
          }
        } else {
          s->block[s->cur_block] = (uint16_t)(s->in_bits >> 10);
          if (s->cur_block == 1 && (s->block[1] & 0x0800)) s->bgroup_off = 4; else s->bgroup_off = 0;
          if (++s->cur_block > 3) {
            s->cur_block = 0; s->block_errors = 0;
            rdssync_emit(s);
          }
        }
      }
      break;
    case ST_GROUPRESYNC:
      if (++s->bit_pos >= 26) {
        s->bit_pos = 0;
        if (++s->cur_block > 3) { s->cur_block = 0; s->state = ST_GROUPDECODE; }
      }
      break;
  }
}

/* ------------------------------------------------------------------------------------------
 * cRDSRxSignalProcessor -- RDSProcess.cpp:43-270
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  float fs, rate;
  float* match_coef; unsigned match_len;
  float last_sync, last_slope, last_data;
  float nco_phase, nco_freq, nco_lo, nco_hi, alpha, beta;
  rfo_fir lp, mf;
  rfo_iir sync;
  rfo_rdsdc dc;
  int last_bit;
  rfo_rdssync bs;
  /* scratch / taps */
  cf32* arr_in; cf32* raw; float* mag; float* data;
  cf32* tap_dec; cf32* tap_lp; float* tap_pll; float* tap_mf; float* tap_sync;
  unsigned cap, nr;
  uint8_t* bits; unsigned nbits, bitcap;
} rdsproc;

static void rdsproc_reset(rdsproc* r)                               /* :90-118 */
{
  r->nco_phase = 0.0f; r->nco_freq = 0.0f;
  fir_init_lp(&r->lp, 0, 1.0f, 40.0f, 2400.0f, (float)(1.3 * 2400.0), r->rate);
  fir_init_const(&r->mf, r->match_len, r->match_coef, r->rate);
  iir_init(&r->sync, 2, (float)(57000.0 / 48.0), 500, r->rate);
  r->last_sync = 0; r->last_slope = 0; r->last_bit = 0;
  rdssync_reset(&r->bs);
  r->last_data = 0;
}

static void rdsproc_init(rdsproc* r, float SampleRate)              /* :43-88 */
{
  memset(r, 0, sizeof(*r));
  r->fs = SampleRate;
  rdsdc_init(&r->dc);
  fir_ctor(&r->lp); fir_ctor(&r->mf);
  r->rate = rdsdc_set_data_rate(&r->dc, r->fs, 8000.0f);
  rdsdc_set_frequency(&r->dc, -57000.0f);
  float norm = (float)(K_2PI / r->rate);
  r->nco_lo = (float)((r->nco_freq - 12.0) * norm);
  r->nco_hi = (float)((r->nco_freq + 12.0) * norm);
  r->alpha = (float)(2.0 * 0.707 * 1.00 * norm);
  r->beta = (float)((r->alpha * r->alpha) / (4.0 * 0.707 * 0.707));
  const double bitrate = 57000.0 / 48.0;
  unsigned L = (unsigned)(r->rate / bitrate);
  r->match_coef = (float*)calloc(L * 2 + 1, sizeof(float));
  for (unsigned i = 0; i <= L; i++) {
    float t = (float)i / r->rate;
    float x = (float)(t * bitrate);
    float x64 = (float)(64.0 * x);
    double v = .75 * cosf((float)(2.0 * K_2PI * x)) * ((1.0 / (1.0 / x - x64)) - (1.0 / (9.0 / x - x64)));
    r->match_coef[i + L] = (float)v;
    r->match_coef[L - i] = (float)(-.75 * cosf((float)(2.0 * K_2PI * x)) * ((1.0 / (1.0 / x - x64)) - (1.0 / (9.0 / x - x64))));
  }
  r->match_len = L * 2;
  rdsproc_reset(r);
}

static void rdsproc_free(rdsproc* r)
{
  rdsdc_delete(&r->dc);
  free(r->match_coef); free(r->arr_in); free(r->raw); free(r->mag); free(r->data);
  free(r->tap_dec); free(r->tap_lp); free(r->tap_pll); free(r->tap_mf); free(r->tap_sync);
  free(r->bits); free(r->bs.groups);
}

static float arctan2_approx(float y, float x)                       /* :187-217 */
{
  if (x == 0.0) {
    if (y > 0.0) return (float)K_PI2;
    if (y == 0.0) return 0.0f;
    return (float)-K_PI2;
  }
  float angle;
  float z = y / x;
  if (fabsf(z) < 1.0) {
    angle = (float)(z / (1.0 + 0.2854 * z * z));
    if (x < 0.0) {
      if (y < 0.0) return (float)(angle - K_PI);
      return (float)(angle + K_PI);
    }
  } else {
    angle = (float)(K_PI2 - z / (z * z + 0.2854));
    if (y < 0.0) return (float)(angle - K_PI);
  }
  return angle;
}

static void rdsproc_pll(rdsproc* r, const cf32* in, float* out, unsigned n) /* :222-270 */
{
  for (unsigned i = 0; i < n; i++) {
    float Sin, Cos;
    rfo_sincos(r->nco_phase, &Sin, &Cos);
    float tre = Cos * in[i].re - Sin * in[i].im;
    float tim = Cos * in[i].im + Sin * in[i].re;
    float phzerror = -arctan2_approx(tim, tre);
    r->nco_freq += (r->beta * phzerror);
    if (r->nco_freq > r->nco_hi) r->nco_freq = r->nco_hi;
    else if (r->nco_freq < r->nco_lo) r->nco_freq = r->nco_lo;
    r->nco_phase += (r->nco_freq + r->alpha * phzerror);
    out[i] = tim;
  }
  r->nco_phase = fmodf(r->nco_phase, (float)K_2PI);
}

static void rdsproc_push_bit(rdsproc* r, int bit)
{
  if (r->nbits == r->bitcap) {
    r->bitcap = r->bitcap ? 2 * r->bitcap : 4096;
    r->bits = (uint8_t*)realloc(r->bits, r->bitcap);
  }
  r->bits[r->nbits++] = (uint8_t)bit;
  rdssync_bit(&r->bs, bit);
}

static void rdsproc_process(rdsproc* r, const float* in, unsigned n) /* :120-180 */
{
  if (n > r->cap) {
    r->cap = n;
    r->arr_in = (cf32*)realloc(r->arr_in, n * sizeof(cf32));
    r->raw = (cf32*)realloc(r->raw, n * sizeof(cf32));
    r->mag = (float*)realloc(r->mag, n * sizeof(float));
    r->data = (float*)realloc(r->data, n * sizeof(float));
    r->tap_dec = (cf32*)realloc(r->tap_dec, n * sizeof(cf32));
    r->tap_lp = (cf32*)realloc(r->tap_lp, n * sizeof(cf32));
    r->tap_pll = (float*)realloc(r->tap_pll, n * sizeof(float));
    r->tap_mf = (float*)realloc(r->tap_mf, n * sizeof(float));
    r->tap_sync = (float*)realloc(r->tap_sync, n * sizeof(float));
  }
  for (unsigned i = 0; i < n; i++) { r->arr_in[i].re = in[i]; r->arr_in[i].im = 0.0f; }
  unsigned len = (unsigned)rdsdc_process(&r->dc, (int)n, r->arr_in, r->raw);
  r->nr = len;
  memcpy(r->tap_dec, r->raw, len * sizeof(cf32));
  fir_process_complex(&r->lp, r->raw, len);
  memcpy(r->tap_lp, r->raw, len * sizeof(cf32));
  rdsproc_pll(r, r->raw, r->data, len);
  memcpy(r->tap_pll, r->data, len * sizeof(float));
  fir_process_real(&r->mf, r->data, len);
  memcpy(r->tap_mf, r->data, len * sizeof(float));
  for (unsigned i = 0; i < len; i++) r->mag[i] = r->data[i] * r->data[i];
  iir_process_real(&r->sync, r->mag, len);
  memcpy(r->tap_sync, r->mag, len * sizeof(float));
  for (unsigned i = 0; i < len; i++) {
    float Data = r->data[i], SyncVal = r->mag[i];
    float Slope = SyncVal - r->last_sync;
    r->last_sync = SyncVal;
    if ((Slope < 0.0) && (r->last_slope * Slope) < 0.0) {
      int bit = (r->last_data >= 0) ? 1 : 0;
      rdsproc_push_bit(r, bit ^ r->last_bit);
      r->last_bit = bit;
    }
    r->last_data = Data;
    r->last_slope = Slope;
  }
}

/* ------------------------------------------------------------------------------------------
 * cFmDecoder -- FmDecode.cpp:237-539
 * ---------------------------------------------------------------------------------------- */
struct rfo_decoder {
  float fs_if, fs_bb, freq_dev, demod_gain;
  int tuning_shift;
  unsigned downsample;
  int stereo;
  float if_level, bb_mean, bb_level;
  finetuner tuner;
  rfo_pilot pilot;
  rfo_downsample rs_in, rs_mono, rs_stereo;
  rdsproc rds;
  rfo_iir notch;
  rfo_fir lp;
  float de_re, de_im, de_alpha;
  float nco_phase, nco_incr, nco_hi, nco_lo, pll_alpha, pll_beta, dc_offset;
  /* buffers / taps */
  unsigned cap, n, nb, na;
  cf32 *in_c, *tuned, *demod;
  float *bb, *mono, *stereo_b, *raw, *pilot38;
  float *tap_mono_rs, *tap_stereo_rs, *tap_lp, *tap_de, *tap_notch;
};

rfo_decoder* rfo_create(double fs_if, double tuning_offset, double fs_pcm, double bw_pcm,
                        unsigned downsample, int usver)
{
  rfo_decoder* d = (rfo_decoder*)calloc(1, sizeof(*d));
  d->fs_if = (float)fs_if;
  d->fs_bb = (float)(fs_if / downsample);
  d->tuning_shift = (int)lrint(-64.0 * tuning_offset / fs_if);
  d->freq_dev = (float)60000.0;
  d->downsample = downsample;
  d->demod_gain = (float)(1.0 / (60000.0 / d->fs_bb * K_2PI));
  finetuner_init(&d->tuner, 64, d->tuning_shift);
  pilot_init(&d->pilot, (float)(19000.0 / d->fs_bb), 50 / d->fs_bb, 0.04f);
  downsample_init(&d->rs_in, 8 * downsample, 0.6 / downsample, downsample, 1);
  downsample_init(&d->rs_mono, (unsigned)(int)(d->fs_bb / 1000.0), bw_pcm / d->fs_bb, d->fs_bb / fs_pcm, 0);
  downsample_init(&d->rs_stereo, (unsigned)(int)(d->fs_bb / 1000.0), bw_pcm / d->fs_bb, d->fs_bb / fs_pcm, 0);
  rdsproc_init(&d->rds, d->fs_bb);
  iir_init(&d->notch, 3, (float)19000.0, 5, (float)fs_pcm);
  fir_ctor(&d->lp);
  fir_init_lp(&d->lp, 0, 1.0f, 60.0f, 15000.0f, (float)(1.4 * 15000.0), (float)fs_pcm);
  {                                                                 /* InitDeemphasis :340-346 */
    float Time = usver ? (float)75E-6 : (float)50E-6, SampleRate = (float)fs_pcm;
    d->de_alpha = (1.0f - expf(-1.0f / (SampleRate * Time)));
    d->de_re = d->de_im = 0.0f;
  }
  float fac = (float)(K_2PI / d->fs_bb);                            /* :305-312 */
  float bandwidth = 0.85f * d->fs_bb;
  float maxFreqDev = 0.95f * (0.5f * d->fs_bb);
  d->nco_lo = (-maxFreqDev) * fac;
  d->nco_hi = (+maxFreqDev) * fac;
  d->pll_alpha = 0.125f * bandwidth * fac;
  d->pll_beta = (d->pll_alpha * d->pll_alpha) / 2.0f;
  rfo_reset(d);
  return d;
}

void rfo_destroy(rfo_decoder* d)
{
  if (!d) return;
  free(d->tuner.table);
  downsample_free(&d->rs_in); downsample_free(&d->rs_mono); downsample_free(&d->rs_stereo);
  rdsproc_free(&d->rds);
  free(d->in_c); free(d->tuned); free(d->demod); free(d->bb); free(d->mono); free(d->stereo_b);
  free(d->raw); free(d->pilot38); free(d->tap_mono_rs); free(d->tap_stereo_rs); free(d->tap_lp);
  free(d->tap_de); free(d->tap_notch);
  free(d);
}

void rfo_reset(rfo_decoder* d)                                      /* :326-338 */
{
  d->stereo = 0; d->if_level = 0; d->bb_mean = 0; d->bb_level = 0;
  d->dc_offset = 0; d->nco_incr = 0.0f; d->nco_phase = 0.0f;
  rdsproc_reset(&d->rds);
}

void rfo_u8_to_cf32(const uint8_t* buf, unsigned n, float* out)     /* RTL_SDR_Source.cpp:207-211 */
{
  for (unsigned i = 0; i < 2 * n; ++i) out[i] = (float)(buf[i] / (255.0 / 2.0) - 1.0);
}

static void demod_pll(rfo_decoder* d, const cf32* sig, float* out, unsigned n) /* :361-415 */
{
  float dc = d->dc_offset;
  for (unsigned i = 0; i < n; ++i) {
    float Sin, Cos;
    rfo_sincos(d->nco_phase, &Sin, &Cos);
    float dre = Cos * sig[i].re - Sin * sig[i].im;
    float dim = Cos * sig[i].im + Sin * sig[i].re;
    float err = -rfo_atan2f(dim, dre);
    d->nco_incr += d->pll_beta * err;
    if (d->nco_incr < d->nco_lo) d->nco_incr = d->nco_lo;
    if (d->nco_incr > d->nco_hi) d->nco_incr = d->nco_hi;
    d->nco_phase += d->nco_incr + d->pll_alpha * err;
    if (d->nco_phase >= K_2PI) d->nco_phase = (float)fmod(d->nco_phase, K_2PI);
    while (d->nco_phase < 0) d->nco_phase = (float)(d->nco_phase + K_2PI);
    float phaseIncr = 2 * d->nco_incr;
    dc = (float)((1 - 0.0001) * dc + 0.0001 * phaseIncr);
    out[i] = (phaseIncr - dc) * d->demod_gain;
  }
  d->dc_offset = dc;
}

static void ensure(rfo_decoder* d, unsigned n)
{
  if (n <= d->cap) return;
  d->cap = n;
#define RA(p, T) d->p = (T*)realloc(d->p, (size_t)n * sizeof(T))
  RA(in_c, cf32); RA(tuned, cf32); RA(demod, cf32); RA(bb, float); RA(mono, float); RA(stereo_b, float);
  RA(raw, float); RA(pilot38, float); RA(tap_mono_rs, float); RA(tap_stereo_rs, float);
#undef RA
  d->tap_lp = (float*)realloc(d->tap_lp, (size_t)2 * n * sizeof(float));
  d->tap_de = (float*)realloc(d->tap_de, (size_t)2 * n * sizeof(float));
  d->tap_notch = (float*)realloc(d->tap_notch, (size_t)2 * n * sizeof(float));
}

unsigned rfo_process_cf32(rfo_decoder* d, const float* iq, unsigned samples, float* audio) /* :417-502 */
{
  ensure(d, samples ? samples : 1);
  d->n = samples;
  unsigned dataSize = samples;
  finetuner_process(&d->tuner, (const cf32*)iq, d->tuned, dataSize);
  {                                                                 /* RMSLevelApprox :505-519 */
    unsigned n = (dataSize + 63) / 64;
    float level = 0;
    for (unsigned i = 0; i < n; ++i) {
      float re = d->tuned[i].re, im = d->tuned[i].im;
      level += re * re + im * im;
    }
    d->if_level = 0.95f * d->if_level + 0.05f * sqrtf(level / n);
  }
  dataSize = downsample_complex(&d->rs_in, d->tuned, d->demod, dataSize);
  d->nb = dataSize;
  demod_pll(d, d->demod, d->bb, dataSize);
  rdsproc_process(&d->rds, d->bb, dataSize);
  {                                                                 /* SamplesMeanRMS :522-539 */
    float vsum = 0, vsumsq = 0;
    for (unsigned i = 0; i < dataSize; ++i) { float v = d->bb[i]; vsum += v; vsumsq += v * v; }
    float mean = vsum / dataSize, rms = sqrtf(vsumsq / dataSize);
    d->bb_mean = 0.95f * d->bb_mean + 0.05f * mean;
    d->bb_level = 0.95f * d->bb_level + 0.05f * rms;
  }
  unsigned monoSize = downsample_real(&d->rs_mono, d->bb, d->mono, dataSize);
  memcpy(d->tap_mono_rs, d->mono, monoSize * sizeof(float));
  d->stereo = pilot_process(&d->pilot, d->bb, d->pilot38, dataSize);
  for (unsigned i = 0; i < dataSize; ++i) d->raw[i] = d->pilot38[i] * (2 * d->bb[i]);
  dataSize = downsample_real(&d->rs_stereo, d->raw, d->stereo_b, dataSize);
  d->na = dataSize;
  memcpy(d->tap_stereo_rs, d->stereo_b, dataSize * sizeof(float));
  fir_process_two(&d->lp, d->stereo_b, d->mono, dataSize);
  memcpy(d->tap_lp, d->stereo_b, dataSize * sizeof(float));
  memcpy(d->tap_lp + dataSize, d->mono, dataSize * sizeof(float));
  for (unsigned i = 0; i < dataSize; ++i) {                         /* deemphasis :348-359 */
    d->de_re = (1.0f - d->de_alpha) * d->de_re + d->de_alpha * d->stereo_b[i];
    d->stereo_b[i] = d->de_re * 2.0f;
    d->de_im = (1.0f - d->de_alpha) * d->de_im + d->de_alpha * d->mono[i];
    d->mono[i] = d->de_im * 2.0f;
  }
  memcpy(d->tap_de, d->stereo_b, dataSize * sizeof(float));
  memcpy(d->tap_de + dataSize, d->mono, dataSize * sizeof(float));
  iir_process_two(&d->notch, d->stereo_b, d->mono, dataSize);
  memcpy(d->tap_notch, d->stereo_b, dataSize * sizeof(float));
  memcpy(d->tap_notch + dataSize, d->mono, dataSize * sizeof(float));
  (void)monoSize;
  if (d->stereo) {                                                  /* :473-499 */
    for (unsigned i = 0; i < dataSize; ++i) {
      float m = d->mono[i], s = d->stereo_b[i];
      audio[2 * i] = (m + s) * 0.5f;
      audio[2 * i + 1] = (m - s) * 0.5f;
    }
  } else {
    for (unsigned i = 0; i < dataSize; ++i) {
      float m = d->mono[i] * 0.5f;
      audio[2 * i] = m;
      audio[2 * i + 1] = m;
    }
  }
  return 2 * dataSize;
}

unsigned rfo_process_u8(rfo_decoder* d, const uint8_t* iq, unsigned n, float* audio)
{
  ensure(d, n ? n : 1);
  rfo_u8_to_cf32(iq, n, (float*)d->in_c);
  return rfo_process_cf32(d, (const float*)d->in_c, n, audio);
}

unsigned rfo_take_groups(rfo_decoder* d, uint16_t* out, unsigned max_groups)
{
  rfo_rdssync* s = &d->rds.bs;
  unsigned n = s->ngroups < max_groups ? s->ngroups : max_groups;
  if (out && n) memcpy(out, s->groups, (size_t)n * 4 * sizeof(uint16_t));
  memmove(s->groups, s->groups + 4 * (size_t)n, (size_t)(s->ngroups - n) * 4 * sizeof(uint16_t));
  s->ngroups -= n;
  return n;
}

unsigned rfo_take_bits(rfo_decoder* d, uint8_t* out, unsigned max_bits)
{
  rdsproc* r = &d->rds;
  unsigned n = r->nbits < max_bits ? r->nbits : max_bits;
  if (out && n) memcpy(out, r->bits, n);
  memmove(r->bits, r->bits + n, r->nbits - n);
  r->nbits -= n;
  return n;
}

void rfo_status(const rfo_decoder* d, float* out)                   /* FmDecode.h:140-165 */
{
  out[0] = d->stereo ? 1.0f : 0.0f;
  out[1] = d->if_level;
  out[2] = d->bb_level;
  out[3] = d->bb_mean;
  out[4] = 2 * d->pilot.level;
  float tuned = -d->tuning_shift * d->fs_if / (float)64;
  out[5] = tuned + d->bb_mean * d->freq_dev;
}

void rfo_constants(const rfo_decoder* d, double* s)
{
  int i = 0;
  s[i++] = d->fs_if; s[i++] = d->fs_bb; s[i++] = d->tuning_shift; s[i++] = d->demod_gain;
  s[i++] = d->nco_lo; s[i++] = d->nco_hi; s[i++] = d->pll_alpha; s[i++] = d->pll_beta;
  s[i++] = d->de_alpha;
  s[i++] = d->pilot.minfreq; s[i++] = d->pilot.maxfreq; s[i++] = d->pilot.b0; s[i++] = d->pilot.a1;
  s[i++] = d->pilot.a2; s[i++] = d->pilot.lb0; s[i++] = d->pilot.lb1; s[i++] = d->pilot.freq;
  s[i++] = d->pilot.minsignal; s[i++] = d->pilot.lock_delay;
  s[i++] = d->rs_in.order; s[i++] = d->rs_in.downsample_int;
  s[i++] = d->rs_mono.order; s[i++] = d->rs_mono.downsample;
  s[i++] = d->rds.rate; s[i++] = d->rds.nco_lo; s[i++] = d->rds.nco_hi; s[i++] = d->rds.alpha;
  s[i++] = d->rds.beta; s[i++] = d->rds.match_len; s[i++] = d->rds.lp.ntaps;
  s[i++] = d->rds.dc.nco_inc; s[i++] = d->rds.dc.osc_cos; s[i++] = d->rds.dc.osc_sin;
  s[i++] = d->rds.sync.A1; s[i++] = d->rds.sync.A2; s[i++] = d->rds.sync.B0; s[i++] = d->rds.sync.B1;
  s[i++] = d->rds.sync.B2;
  s[i++] = d->notch.A1; s[i++] = d->notch.A2; s[i++] = d->notch.B0; s[i++] = d->notch.B1; s[i++] = d->notch.B2;
  s[i++] = d->lp.ntaps;
  s[i++] = d->rds.dc.nstages;
  for (int k = 0; k < 6; ++k) s[i++] = k < d->rds.dc.nstages ? d->rds.dc.st[k].len : 0;
}

unsigned rfo_table(const rfo_decoder* d, int which, float* out, unsigned max_floats)
{
  const float* src = NULL; unsigned n = 0;
  switch (which) {
    case 0: src = (const float*)d->tuner.table; n = 2 * d->tuner.size; break;
    case 1: src = d->rs_in.coeff; n = d->rs_in.order + 2; break;
    case 2: src = d->rs_mono.coeff; n = d->rs_mono.order + 2; break;
    case 3: src = d->rds.lp.coef; n = d->rds.lp.ntaps; break;
    case 4: src = d->rds.mf.coef; n = d->rds.mf.ntaps; break;
    case 5: src = d->lp.coef; n = d->lp.ntaps; break;
    default: return 0;
  }
  if (n > max_floats) n = max_floats;
  if (out) memcpy(out, src, n * sizeof(float));
  return n;
}

const float* rfo_tap(const rfo_decoder* d, const char* name, unsigned* nf)
{
#define T(nm, ptr, cnt) if (!strcmp(name, nm)) { *nf = (cnt); return (const float*)(ptr); }
  T("tuned", d->tuned, 2 * d->n) T("demod_in", d->demod, 2 * d->nb) T("baseband", d->bb, d->nb)
  T("rds_dec", d->rds.tap_dec, 2 * d->rds.nr) T("rds_lp", d->rds.tap_lp, 2 * d->rds.nr)
  T("rds_pll", d->rds.tap_pll, d->rds.nr) T("rds_mf", d->rds.tap_mf, d->rds.nr)
  T("rds_sync", d->rds.tap_sync, d->rds.nr) T("mono_rs", d->tap_mono_rs, d->na)
  T("pilot38", d->pilot38, d->nb) T("rawstereo", d->raw, d->nb) T("stereo_rs", d->tap_stereo_rs, d->na)
  T("lp", d->tap_lp, 2 * d->na) T("deemph", d->tap_de, 2 * d->na) T("notch", d->tap_notch, 2 * d->na)
#undef T
  *nf = 0;
  return NULL;
}

unsigned rfo_last_stereo(const rfo_decoder* d) { return (unsigned)d->stereo; }

/* ------------------------------------------------------------------------------------------
 * stand-alone primitive wrappers
 * ---------------------------------------------------------------------------------------- */
struct rfo_freqshift { float nco_freq, nco_inc, nco_time, in_rate; };

rfo_freqshift* rfo_freqshift_create(float nco_freq, float in_rate)  /* FreqShift.cpp:10-16 */
{
  rfo_freqshift* f = (rfo_freqshift*)calloc(1, sizeof(*f));
  f->nco_time = 0.0f; f->in_rate = in_rate; f->nco_freq = nco_freq;
  f->nco_inc = (float)(K_2PI * f->nco_freq / f->in_rate);
  return f;
}
void rfo_freqshift_destroy(rfo_freqshift* f) { free(f); }
void rfo_freqshift_reset(rfo_freqshift* f) { f->nco_time = 0.0f; }
void rfo_freqshift_process(rfo_freqshift* f, float* iq, unsigned n) /* :23-76, x86 branch: no wrap */
{
  cf32* x = (cf32*)iq;
  float acc = f->nco_time;
  for (unsigned i = 0; i < n; ++i) {
    float s, c;
    rfo_sincos(acc, &s, &c);
    acc += f->nco_inc;
    cf32 d = x[i];
    x[i].re = (d.re * c) - (d.im * s);
    x[i].im = (d.re * s) + (d.im * c);
  }
  f->nco_time = acc;
}

rfo_downsample* rfo_downsample_create(unsigned order, double cutoff, double downsample, int integer_factor)
{
  rfo_downsample* f = (rfo_downsample*)calloc(1, sizeof(*f));
  downsample_init(f, order, cutoff, downsample, integer_factor);
  return f;
}
void rfo_downsample_destroy(rfo_downsample* f) { if (f) { downsample_free(f); free(f); } }
void rfo_downsample_reset(rfo_downsample* f)
{
  f->pos_int = 0; f->pos_frac = 0;
  memset(f->state_r, 0, f->order * sizeof(float));
  memset(f->state_c, 0, f->order * sizeof(cf32));
}
unsigned rfo_downsample_process_real(rfo_downsample* f, const float* in, float* out, unsigned n)
{ return downsample_real(f, in, out, n); }
unsigned rfo_downsample_process_complex(rfo_downsample* f, const float* in, float* out, unsigned n)
{ return downsample_complex(f, (const cf32*)in, (cf32*)out, n); }
unsigned rfo_downsample_coeff(const rfo_downsample* f, float* out)
{ if (out) memcpy(out, f->coeff, (f->order + 2) * sizeof(float)); return f->order + 2; }

rfo_rdsdc* rfo_rdsdc_create(void) { rfo_rdsdc* d = (rfo_rdsdc*)malloc(sizeof(*d)); rdsdc_init(d); return d; }
void rfo_rdsdc_destroy(rfo_rdsdc* d) { if (d) { rdsdc_delete(d); free(d); } }
void rfo_rdsdc_set_frequency(rfo_rdsdc* d, float f) { rdsdc_set_frequency(d, f); }
float rfo_rdsdc_set_data_rate(rfo_rdsdc* d, float r, float bw) { return rdsdc_set_data_rate(d, r, bw); }
float rfo_rdsdc_set_wfm_data_rate(rfo_rdsdc* d, float r, float bw) { return rdsdc_set_wfm_data_rate(d, r, bw); }
int rfo_rdsdc_process(rfo_rdsdc* d, int n, float* inout, float* out) { return rdsdc_process(d, n, (cf32*)inout, (cf32*)out); }
int rfo_rdsdc_stages(const rfo_rdsdc* d, int* lens, int max)
{ for (int i = 0; i < d->nstages && i < max; ++i) lens[i] = d->st[i].len; return d->nstages; }

rfo_fir* rfo_fir_create(void) { rfo_fir* f = (rfo_fir*)malloc(sizeof(*f)); fir_ctor(f); return f; }
void rfo_fir_destroy(rfo_fir* f) { free(f); }
int rfo_fir_init_lp(rfo_fir* f, unsigned taps, float scale, float astop, float fpass, float fstop, float fs)
{ return fir_init_lp(f, taps, scale, astop, fpass, fstop, fs); }
int rfo_fir_init_hp(rfo_fir* f, unsigned taps, float scale, float astop, float fpass, float fstop, float fs)
{ return fir_init_hp(f, taps, scale, astop, fpass, fstop, fs); }
void rfo_fir_init_const(rfo_fir* f, unsigned taps, const float* coef, float fs)
{
  fir_init_const(f, taps, coef, fs);
}
unsigned rfo_fir_coef(const rfo_fir* f, float* out) { if (out) memcpy(out, f->coef, f->ntaps * sizeof(float)); return f->ntaps; }
void rfo_fir_process_real(rfo_fir* f, float* buf, unsigned n) { fir_process_real(f, buf, n); }
void rfo_fir_process_complex(rfo_fir* f, float* buf, unsigned n) { fir_process_complex(f, (cf32*)buf, n); }
void rfo_fir_process_two(rfo_fir* f, float* a, float* b, unsigned n) { fir_process_two(f, a, b, n); }

rfo_iir* rfo_iir_create(void) { return (rfo_iir*)calloc(1, sizeof(rfo_iir)); }
void rfo_iir_destroy(rfo_iir* f) { free(f); }
int rfo_iir_init(rfo_iir* f, int type, float f0, float q, float fs) { return iir_init(f, type, f0, q, fs); }
void rfo_iir_coef(const rfo_iir* f, float* out) { out[0] = f->A1; out[1] = f->A2; out[2] = f->B0; out[3] = f->B1; out[4] = f->B2; }
void rfo_iir_process_real(rfo_iir* f, float* buf, unsigned n) { iir_process_real(f, buf, n); }
void rfo_iir_process_complex(rfo_iir* f, float* buf, unsigned n)
{
  cf32* x = (cf32*)buf;
  for (unsigned i = 0; i < n; ++i) {
    x[i].re = biquad_step(f, x[i].re, &f->w1a, &f->w2a);
    x[i].im = biquad_step(f, x[i].im, &f->w1b, &f->w2b);
  }
}
void rfo_iir_process_two(rfo_iir* f, float* a, float* b, unsigned n) { iir_process_two(f, a, b, n); }

rfo_pilot* rfo_pilot_create(float freq, float bw, float minsig)
{ rfo_pilot* p = (rfo_pilot*)calloc(1, sizeof(*p)); pilot_init(p, freq, bw, minsig); return p; }
void rfo_pilot_destroy(rfo_pilot* p) { free(p); }
int rfo_pilot_process(rfo_pilot* p, const float* in, float* out, unsigned n) { return pilot_process(p, in, out, n); }
float rfo_pilot_level(const rfo_pilot* p) { return 2 * p->level; }

rfo_rdssync* rfo_rdssync_create(void) { rfo_rdssync* s = (rfo_rdssync*)calloc(1, sizeof(*s)); rdssync_reset(s); return s; }
void rfo_rdssync_destroy(rfo_rdssync* s) { if (s) { free(s->groups); free(s); } }
void rfo_rdssync_reset(rfo_rdssync* s) { rdssync_reset(s); }
void rfo_rdssync_push_bits(rfo_rdssync* s, const uint8_t* bits, unsigned n)
{ for (unsigned i = 0; i < n; ++i) rdssync_bit(s, bits[i] & 1); }
unsigned rfo_rdssync_take_groups(rfo_rdssync* s, uint16_t* out, unsigned max_groups)
{
  unsigned n = s->ngroups < max_groups ? s->ngroups : max_groups;
  if (out && n) memcpy(out, s->groups, (size_t)n * 4 * sizeof(uint16_t));
  memmove(s->groups, s->groups + 4 * (size_t)n, (size_t)(s->ngroups - n) * 4 * sizeof(uint16_t));
  s->ngroups -= n;
  return n;
}
