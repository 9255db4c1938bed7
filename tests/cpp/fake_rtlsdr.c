/* tests/cpp/fake_rtlsdr.c -- a stand-in for librtlsdr.so (test infrastructure): the fifteen entry points the source
 * adapter binds (RTL_SDR_Source.cpp), a "device" that plays a file.
 *   FAKE_RTLSDR_FILE   raw u8 IQ; rtlsdr_read_async hands it out in buffers of buf_len bytes, inserts ONE short
 *                      buffer (buf_len / 2) after the first full one, then waits for rtlsdr_cancel_async
 *   FAKE_RTLSDR_FAIL   n: the first n rtlsdr_read_async calls return -1 at once (restart logic)
 *   FAKE_RTLSDR_LOG    every configuration call is appended to this file, one line each */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef void (*read_cb_t)(unsigned char*, uint32_t, void*);
typedef struct
{
  uint32_t rate, freq;
  int gain, gain_mode, agc;
  volatile int cancel;
  int reads;
} fake_dev;

static void logf_(const char* fmt, long a, long b)
{
  const char* p = getenv("FAKE_RTLSDR_LOG");
  if (!p)
    return;
  FILE* f = fopen(p, "a");
  if (!f)
    return;
  fprintf(f, fmt, a, b);
  fputc('\n', f);
  fclose(f);
}

uint32_t rtlsdr_get_device_count(void) { return 1; }
const char* rtlsdr_get_device_name(uint32_t i) { return i == 0 ? "Fake RTL2838 (file player)" : ""; }
int rtlsdr_open(void** dev, uint32_t index)
{
  if (index != 0)
    return -1;
  *dev = calloc(1, sizeof(fake_dev));
  logf_("open %ld", (long)index, 0);
  return 0;
}
int rtlsdr_close(void* d)
{
  logf_("close", 0, 0);
  free(d);
  return 0;
}
int rtlsdr_set_sample_rate(void* d, uint32_t r) { ((fake_dev*)d)->rate = r; logf_("sample_rate %ld", (long)r, 0); return 0; }
uint32_t rtlsdr_get_sample_rate(void* d) { return ((fake_dev*)d)->rate; }
int rtlsdr_set_center_freq(void* d, uint32_t f) { ((fake_dev*)d)->freq = f; logf_("center_freq %ld", (long)f, 0); return 0; }
uint32_t rtlsdr_get_center_freq(void* d) { return ((fake_dev*)d)->freq; }
int rtlsdr_set_tuner_gain_mode(void* d, int m) { ((fake_dev*)d)->gain_mode = m; logf_("gain_mode %ld", m, 0); return 0; }
int rtlsdr_set_tuner_gain(void* d, int g) { ((fake_dev*)d)->gain = g; logf_("gain %ld", g, 0); return 0; }
int rtlsdr_get_tuner_gain(void* d) { return ((fake_dev*)d)->gain; }
int rtlsdr_set_agc_mode(void* d, int on) { ((fake_dev*)d)->agc = on; logf_("agc %ld", on, 0); return 0; }
int rtlsdr_reset_buffer(void* d) { (void)d; logf_("reset_buffer", 0, 0); return 0; }
int rtlsdr_cancel_async(void* d) { ((fake_dev*)d)->cancel = 1; return 0; }

int rtlsdr_read_async(void* dv, read_cb_t cb, void* ctx, uint32_t buf_num, uint32_t buf_len)
{
  fake_dev* d = (fake_dev*)dv;
  const char* fail = getenv("FAKE_RTLSDR_FAIL");
  logf_("read_async %ld %ld", (long)buf_num, (long)buf_len);
  if (fail && d->reads++ < atoi(fail))
    return -1;
  const char* path = getenv("FAKE_RTLSDR_FILE");
  FILE* f = path ? fopen(path, "rb") : NULL;
  unsigned char* buf = (unsigned char*)malloc(buf_len);
  int n = 0;
  while (f && !d->cancel && fread(buf, 1, buf_len, f) == buf_len)
  {
    cb(buf, buf_len, ctx);
    if (n++ == 0)
      cb(buf, buf_len / 2, ctx); /* a short read: must be dropped and counted */
  }
  if (f)
    fclose(f);
  free(buf);
  while (!d->cancel)
    usleep(2000);
  return 0;
}
