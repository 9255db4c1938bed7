/* oracle/radiofm_oracle.h -- plain-C restatement of the reference IQ->audio(+RDS) chain.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (pvr.rtl.radiofm_b200/, include/) may
 * include, link or call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg do.  Parity pin: tests/test_oracle_port.py checks every stage of this restatement BIT FOR
 * BIT against the unmodified reference compiled by oracle/Makefile (oracle/_ref/
 * libradiofm_ref.so) and against the fixtures under tests/golden/ generated from that library
 * (tests/golden/make_golden.py).  The reference itself ships no tests or golden vectors
 * (SURVEY.md section 4).
 */
#ifndef RADIOFM_ORACLE_H
#define RADIOFM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rfo_decoder rfo_decoder;

/* cFmDecoder ctor, FmDecode.cpp:237-314 */
rfo_decoder* rfo_create(double fs_if, double tuning_offset, double fs_pcm, double bw_pcm,
                        unsigned downsample, int usver);
void rfo_destroy(rfo_decoder* d);
/* cFmDecoder::Reset, FmDecode.cpp:326-338 */
void rfo_reset(rfo_decoder* d);
/* RTL_SDR_Source.cpp:207-211 */
void rfo_u8_to_cf32(const uint8_t* iq, unsigned n, float* out);
/* cFmDecoder::ProcessStream, FmDecode.cpp:417-502; returns floats written (2 x frames) */
unsigned rfo_process_cf32(rfo_decoder* d, const float* iq, unsigned n, float* audio);
unsigned rfo_process_u8(rfo_decoder* d, const uint8_t* iq, unsigned n, float* audio);

unsigned rfo_take_groups(rfo_decoder* d, uint16_t* out, unsigned max_groups);
unsigned rfo_take_bits(rfo_decoder* d, uint8_t* out, unsigned max_bits);
/* out[0]=stereo [1]=IF level [2]=bb level [3]=bb mean [4]=pilot level [5]=tuning offset */
void rfo_status(const rfo_decoder* d, float* out);
/* same index list as ref_fm_constants() in oracle/ref_harness.cpp */
void rfo_constants(const rfo_decoder* d, double* s);
/* same `which` as ref_fm_table() */
unsigned rfo_table(const rfo_decoder* d, int which, float* out, unsigned max_floats);

/* Stage taps of the LAST process call (pointers into decoder-owned memory).
 * names: tuned demod_in baseband rds_dec rds_lp rds_pll rds_mf rds_sync mono_rs pilot38
 *        rawstereo stereo_rs lp deemph notch   (lp/deemph/notch: stereo[na] then mono[na]) */
const float* rfo_tap(const rfo_decoder* d, const char* name, unsigned* n_floats);
unsigned rfo_last_stereo(const rfo_decoder* d);

/* --- stand-alone primitives (same semantics as the reference classes) --- */
float rfo_atan2f(float y, float x);            /* glibc 2.39 atan2f restated (fdlibm float) */
void rfo_sincos(float phase, float* s, float* c); /* x87 fsincos on a float, rounded to float */

typedef struct rfo_freqshift rfo_freqshift;    /* cFreqShift, FreqShift.cpp:10-76 (x86 branch) */
rfo_freqshift* rfo_freqshift_create(float nco_freq, float in_rate);
void rfo_freqshift_destroy(rfo_freqshift* f);
void rfo_freqshift_reset(rfo_freqshift* f);
void rfo_freqshift_process(rfo_freqshift* f, float* iq, unsigned n);

typedef struct rfo_downsample rfo_downsample;  /* cDownsampleFilter, DownConvert.cpp:18-256 */
rfo_downsample* rfo_downsample_create(unsigned order, double cutoff, double downsample, int integer_factor);
void rfo_downsample_destroy(rfo_downsample* f);
void rfo_downsample_reset(rfo_downsample* f);
unsigned rfo_downsample_process_real(rfo_downsample* f, const float* in, float* out, unsigned n);
unsigned rfo_downsample_process_complex(rfo_downsample* f, const float* in, float* out, unsigned n);
unsigned rfo_downsample_coeff(const rfo_downsample* f, float* out);

typedef struct rfo_rdsdc rfo_rdsdc;            /* CRDSDownConvert, DownConvert.cpp:271-727 */
rfo_rdsdc* rfo_rdsdc_create(void);
void rfo_rdsdc_destroy(rfo_rdsdc* d);
void rfo_rdsdc_set_frequency(rfo_rdsdc* d, float f);
float rfo_rdsdc_set_data_rate(rfo_rdsdc* d, float in_rate, float max_bw);
float rfo_rdsdc_set_wfm_data_rate(rfo_rdsdc* d, float in_rate, float max_bw);
int rfo_rdsdc_process(rfo_rdsdc* d, int n, float* inout, float* out);
int rfo_rdsdc_stages(const rfo_rdsdc* d, int* lens, int max);

typedef struct rfo_fir rfo_fir;                /* cFirFilter, FirFilter.cpp */
rfo_fir* rfo_fir_create(void);
void rfo_fir_destroy(rfo_fir* f);
int rfo_fir_init_lp(rfo_fir* f, unsigned taps, float scale, float astop, float fpass, float fstop, float fs);
int rfo_fir_init_hp(rfo_fir* f, unsigned taps, float scale, float astop, float fpass, float fstop, float fs); /* :195-264 */
void rfo_fir_init_const(rfo_fir* f, unsigned taps, const float* coef, float fs);
unsigned rfo_fir_coef(const rfo_fir* f, float* out);
void rfo_fir_process_real(rfo_fir* f, float* buf, unsigned n);
void rfo_fir_process_complex(rfo_fir* f, float* buf, unsigned n);
void rfo_fir_process_two(rfo_fir* f, float* a, float* b, unsigned n);

typedef struct rfo_iir rfo_iir;                /* cIirFilter, IirFilter.cpp */
rfo_iir* rfo_iir_create(void);
void rfo_iir_destroy(rfo_iir* f);
int rfo_iir_init(rfo_iir* f, int type, float f0, float q, float fs);
void rfo_iir_coef(const rfo_iir* f, float* out);
void rfo_iir_process_real(rfo_iir* f, float* buf, unsigned n);
void rfo_iir_process_complex(rfo_iir* f, float* buf, unsigned n);
void rfo_iir_process_two(rfo_iir* f, float* a, float* b, unsigned n);

typedef struct rfo_pilot rfo_pilot;            /* cPilotPhaseLock, FmDecode.cpp:88-229 */
rfo_pilot* rfo_pilot_create(float freq, float bandwidth, float minsignal);
void rfo_pilot_destroy(rfo_pilot* p);
int rfo_pilot_process(rfo_pilot* p, const float* in, float* out, unsigned n);
float rfo_pilot_level(const rfo_pilot* p);

/* RDS block sync / FEC (RDSProcess.cpp:272-431), integer only */
typedef struct rfo_rdssync rfo_rdssync;
rfo_rdssync* rfo_rdssync_create(void);
void rfo_rdssync_destroy(rfo_rdssync* s);
void rfo_rdssync_reset(rfo_rdssync* s);
void rfo_rdssync_push_bits(rfo_rdssync* s, const uint8_t* bits, unsigned n);
unsigned rfo_rdssync_take_groups(rfo_rdssync* s, uint16_t* out, unsigned max_groups);
uint32_t rfo_rds_check_block(uint32_t word26, uint32_t offset_syndrome, int use_fec, uint32_t* corrected);

#ifdef __cplusplus
}
#endif
#endif
