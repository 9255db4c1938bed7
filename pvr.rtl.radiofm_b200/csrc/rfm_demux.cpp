// rfm_demux.cpp -- the caller of the hot path: the IQ block queue between the source
// thread and the demux thread (cRadioReceiver::WriteDataBuffer / EndDataBuffer / SourceGetSamples / SourceQueuedSamples,
// RadioReceiver.cpp:420-460) and the packetiser cRadioReceiver::DemuxRead (RadioReceiver.cpp:462-542): stream-change
// packet, UECP packets on stream id 2, audio packets on stream id 1 with their pts / duration, the audio level meter
// (:528-529, :584-598).  Host code over the C ABI of the decoder; the Kodi packet allocation is replaced by a packet
// view into buffers this object owns.
//
// Differences from the reference, all in the plumbing:
//   * blocks stay u8 until they are on the device (the reference queues std::vector<ComplexType>, 4x the bytes) and
//     live in pinned host memory, so the H2D copy of a block is asynchronous;
//   * read-ahead: when an audio packet is handed out and the queue already holds the next block, that block is
//     submitted at once (rfm_decoder_submit_u8) and decodes while the caller consumes the packet.  The packet
//     ORDER is the reference's: the UECP frames a block produced are offered before the next block's audio.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <atomic>
#include <deque>
#include <mutex>
#include <vector>

#include "../../include/radiofm_b200.h"

namespace
{
constexpr double kStreamTimeBase = 1000000.0; // STREAM_TIME_BASE (Kodi DVD_TIME_BASE)

struct PinnedBlock
{
  uint8_t* p = nullptr;
  uint32_t n = 0; // IQ samples
};
} // namespace

struct rfm_demux
{
  rfm_decoder* dec = nullptr;
  rfm_config cfg;
  // ---- source side (any thread)
  std::mutex mu;
  std::condition_variable cv;
  std::deque<PinnedBlock> queue;
  std::vector<PinnedBlock> free_blocks;
  size_t queued_samples = 0;
  bool end_marked = false;
  // ---- demux side (one thread)
  std::atomic<bool> stream_change{true}; // OpenLiveStream sets it, RadioReceiver.cpp:345 (any thread: SetStreamChange)
  double pts_next = kStreamTimeBase;  // :347
  float audio_level = 0.0f;           // :188
  std::vector<uint8_t> uecp;          // m_UECPOutputBuffer
  std::vector<uint8_t> uecp_packet;
  float* audio[2] = {nullptr, nullptr}; // pinned, alternate per block
  uint32_t audio_cap = 0;
  int cur = 0;
  bool in_flight = false;             // a block has been submitted and not yet collected
  PinnedBlock flying;
  uint32_t flying_floats = 0;
  uint64_t blocks_done = 0;
  uint32_t source_block = 0;   // cRtlSdrSource::m_BlockLength for rfm_demux_source_cb
  uint64_t short_reads = 0;
  int pending_error = RFM_OK;  // a read-ahead submit failed after a packet had been completed: reported by the next read
  // what GetSignalStatus reports (RadioReceiver.cpp:544-556), cached under `mu` when a block is collected: the status
  // call comes from another thread and must neither touch the decoder nor wait for the block in flight
  float st_interface_level = 0.0f, st_audio_level = 0.0f;
  int st_stereo = 0;
  bool st_valid = false;
};

namespace
{

void Recycle(rfm_demux* m, const PinnedBlock& b);

// The block goes into the OTHER audio buffer; `cur` only flips once the decoder has accepted it, and a refused block's
// pinned memory goes back to the pool.
int Submit(rfm_demux* m, const PinnedBlock& b)
{
  uint32_t nfl = 0;
  const int rc = rfm_decoder_submit_u8(m->dec, b.p, b.n, m->audio[m->cur ^ 1], m->audio_cap, &nfl);
  if (rc != RFM_OK)
  {
    Recycle(m, b);
    return rc;
  }
  m->cur ^= 1;
  m->flying = b;
  m->flying_floats = nfl;
  m->in_flight = true;
  return RFM_OK;
}

// SourceGetSamples, RadioReceiver.cpp:445-460: waits until a block or the end mark arrives
bool PopBlock(rfm_demux* m, PinnedBlock* out, bool wait)
{
  std::unique_lock<std::mutex> lock(m->mu);
  while (wait && m->queue.empty() && !m->end_marked)
    m->cv.wait_for(lock, std::chrono::milliseconds(20));
  if (m->queue.empty())
    return false;
  *out = m->queue.front();
  m->queue.pop_front();
  m->queued_samples -= out->n;
  return true;
}

void Recycle(rfm_demux* m, const PinnedBlock& b)
{
  std::lock_guard<std::mutex> lock(m->mu);
  m->free_blocks.push_back(b);
}

} // namespace

extern "C"
{

int rfm_demux_create(const rfm_config* cfg, rfm_demux** out)
{
  if (!cfg || !out)
    return RFM_ERR_INVALID;
  *out = nullptr;
  rfm_config c = *cfg;
  c.n_streams = 1; // one tuner, one programme: the add-on's case
  c.n_groups = 1;
  rfm_decoder* dec = nullptr;
  int rc = rfm_decoder_create(&c, &dec);
  if (rc != RFM_OK)
    return rc;
  rfm_demux* m = new rfm_demux();
  m->dec = dec;
  m->cfg = c;
  m->audio_cap = rfm_decoder_max_audio_floats(dec, c.max_block_len);
  m->audio_cap += m->audio_cap & 1u;
  for (int i = 0; i < 2; ++i)
    if (cudaHostAlloc(reinterpret_cast<void**>(&m->audio[i]), (size_t)m->audio_cap * sizeof(float), cudaHostAllocDefault) != cudaSuccess)
    {
      rfm_decoder_destroy(dec);
      delete m;
      return RFM_ERR_CUDA;
    }
  *out = m;
  return RFM_OK;
}

void rfm_demux_destroy(rfm_demux* m)
{
  if (!m)
    return;
  rfm_decoder_synchronize(m->dec);
  rfm_decoder_destroy(m->dec);
  for (auto& b : m->queue)
    cudaFreeHost(b.p);
  for (auto& b : m->free_blocks)
    cudaFreeHost(b.p);
  if (m->in_flight)
    cudaFreeHost(m->flying.p);
  for (float* a : m->audio)
    if (a)
      cudaFreeHost(a);
  delete m;
}

rfm_decoder* rfm_demux_decoder(rfm_demux* m) { return m ? m->dec : nullptr; }

/* WriteDataBuffer, RadioReceiver.cpp:426-436 (source thread) */
int rfm_demux_write_u8(rfm_demux* m, const uint8_t* iq, uint32_t n)
{
  if (!m || (!iq && n))
    return RFM_ERR_INVALID;
  if (n == 0)
    return RFM_OK;
  if (n > m->cfg.max_block_len)
    return RFM_ERR_INVALID;
  PinnedBlock b;
  {
    std::lock_guard<std::mutex> lock(m->mu);
    if (!m->free_blocks.empty())
    {
      b = m->free_blocks.back();
      m->free_blocks.pop_back();
    }
  }
  if (!b.p && cudaHostAlloc(reinterpret_cast<void**>(&b.p), (size_t)m->cfg.max_block_len * 2, cudaHostAllocDefault) != cudaSuccess)
    return RFM_ERR_CUDA;
  memcpy(b.p, iq, (size_t)n * 2);
  b.n = n;
  std::lock_guard<std::mutex> lock(m->mu);
  m->queued_samples += n;
  m->queue.push_back(b);
  m->cv.notify_one();
  return RFM_OK;
}

/* EndDataBuffer, RadioReceiver.cpp:438-443 */
int rfm_demux_end(rfm_demux* m)
{
  if (!m)
    return RFM_ERR_INVALID;
  std::lock_guard<std::mutex> lock(m->mu);
  m->end_marked = true;
  m->cv.notify_all();
  return RFM_OK;
}

/* SourceQueuedSamples, RadioReceiver.cpp:420-424 */
uint64_t rfm_demux_queued_samples(rfm_demux* m)
{
  if (!m)
    return 0;
  std::lock_guard<std::mutex> lock(m->mu);
  return m->queued_samples;
}

/* SetStreamChange, RadioReceiver.h:83 */
void rfm_demux_set_stream_change(rfm_demux* m)
{
  if (m)
    m->stream_change = true;
}

/* DemuxRead, RadioReceiver.cpp:462-542.  Returns RFM_OK with a packet, RFM_DEMUX_END when the end was marked and
 * every block has been handed out (the reference returns nullptr there). */
int rfm_demux_read(rfm_demux* m, rfm_demux_packet* pkt)
{
  if (!m || !pkt)
    return RFM_ERR_INVALID;
  memset(pkt, 0, sizeof(*pkt));
  if (m->pending_error != RFM_OK)
  {
    const int rc = m->pending_error;
    m->pending_error = RFM_OK;
    return rc;
  }
  if (m->stream_change)
  {
    pkt->stream_id = RFM_DEMUX_STREAMCHANGE;
    m->stream_change = false;
    return RFM_OK;
  }
  if (!m->uecp.empty())
  {
    m->uecp_packet.swap(m->uecp);
    m->uecp.clear();
    pkt->stream_id = 2;
    pkt->data = m->uecp_packet.data();
    pkt->size_bytes = (uint32_t)m->uecp_packet.size();
    pkt->pts = m->pts_next;
    return RFM_OK;
  }
  if (!m->in_flight)
  {
    PinnedBlock b;
    if (!PopBlock(m, &b, true))
      return RFM_DEMUX_END;
    const int rc = Submit(m, b);
    if (rc != RFM_OK)
      return rc;
  }
  int rc = rfm_decoder_synchronize(m->dec);
  if (rc != RFM_OK)
    return rc;
  m->in_flight = false;
  Recycle(m, m->flying);
  const float* a = m->audio[m->cur];
  const uint32_t nfl = m->flying_floats;
  // the frames the group decoder produced while this block was decoded (m_UECPOutputBuffer)
  {
    uint8_t tmp[4096];
    uint32_t k = 0;
    do
    {
      rc = rfm_decoder_rds_take_uecp(m->dec, 0, tmp, sizeof(tmp), &k);
      if (rc != RFM_OK)
        return rc;
      m->uecp.insert(m->uecp.end(), tmp, tmp + k);
    } while (k == sizeof(tmp));
  }
  // audio level, RadioReceiver.cpp:528-529 with SamplesMeanRMS (:584-598): float sums, double square root
  {
    float vsum = 0.0f, vsumsq = 0.0f;
    for (uint32_t i = 0; i < nfl; ++i)
    {
      const float v = a[i];
      vsum += v;
      vsumsq += v * v;
    }
    const double rms = sqrt((double)(vsumsq / (float)nfl));
    (void)vsum;
    m->audio_level = (float)(0.95 * m->audio_level + 0.05 * rms);
  }
  {
    rfm_stream_status st;
    rc = rfm_decoder_get_status(m->dec, 0, &st); // the decoder is idle here: the next block is submitted further down
    if (rc != RFM_OK)
      return rc;
    std::lock_guard<std::mutex> lock(m->mu);
    m->st_interface_level = st.interface_level;
    m->st_audio_level = m->audio_level;
    m->st_stereo = st.stereo_detected;
    m->st_valid = true;
  }
  {
    // this loop only consumes the UECP stream: the decoded groups themselves would pile up in the block synchroniser
    uint32_t k = 0;
    do
    {
      rc = rfm_decoder_rds_take_groups(m->dec, 0, nullptr, 1024, &k);
    } while (rc == RFM_OK && k == 1024);
    if (rc != RFM_OK)
      return rc;
  }
  const double duration = (double)nfl * kStreamTimeBase / 2 / m->cfg.sample_rate_pcm;
  pkt->stream_id = 1;
  pkt->data = a;
  pkt->size_bytes = nfl * (uint32_t)sizeof(float);
  pkt->duration = duration;
  pkt->pts = m->pts_next;
  m->pts_next = m->pts_next + duration;
  m->blocks_done += 1;
  // read-ahead: the next block, if it is already queued, decodes while the caller consumes this packet
  PinnedBlock nb;
  if (PopBlock(m, &nb, false))
    m->pending_error = Submit(m, nb); // this packet is complete and is handed out; a failure surfaces on the next read
  // (this packet's audio buffer is the OTHER one: it stays valid until the next audio packet is read)
  return RFM_OK;
}

/* cRtlSdrSource::Configure's block-length rule, RTL_SDR_Source.cpp:124-126 */
uint32_t rfm_source_block_length(uint32_t requested)
{
  uint32_t n = requested < 4096u ? 4096u : (requested > 1024u * 1024u ? 1024u * 1024u : requested);
  return n - n % 4096u;
}

int rfm_demux_set_source_block_length(rfm_demux* m, uint32_t requested)
{
  if (!m)
    return RFM_ERR_INVALID;
  const uint32_t n = rfm_source_block_length(requested);
  if (n > m->cfg.max_block_len)
    return RFM_ERR_INVALID;
  m->source_block = n;
  return RFM_OK;
}

/* cRtlSdrSource::ReadAsyncCB, RTL_SDR_Source.cpp:196-213: the signature of rtlsdr_read_async_cb_t with ctx = the
 * rfm_demux.  A buffer that is not exactly one block is dropped ("short read, samples lost"); a good one is queued as
 * it is -- the u8 -> float conversion of :207-211 happens on the device. */
void rfm_demux_source_cb(unsigned char* buf, uint32_t len, void* ctx)
{
  rfm_demux* m = static_cast<rfm_demux*>(ctx);
  if (!m || !buf)
    return;
  if (m->source_block == 0 || len != 2 * m->source_block)
  {
    std::lock_guard<std::mutex> lock(m->mu);
    m->short_reads += 1;
    return;
  }
  rfm_demux_write_u8(m, buf, m->source_block);
}

uint64_t rfm_demux_short_reads(rfm_demux* m)
{
  if (!m)
    return 0;
  std::lock_guard<std::mutex> lock(m->mu);
  return m->short_reads;
}

float rfm_demux_audio_level(const rfm_demux* m)
{
  if (!m)
    return 0.0f;
  rfm_demux* mm = const_cast<rfm_demux*>(m);
  std::lock_guard<std::mutex> lock(mm->mu);
  return mm->st_valid ? mm->st_audio_level : 0.0f; // the value cached when the last block was collected (any thread)
}

/* GetSignalStatus, RadioReceiver.cpp:544-556: interface level and audio level in dB, stereo flag */
int rfm_demux_signal_status(rfm_demux* m, float* interface_level, float* audio_level_db, int* stereo)
{
  if (!m)
    return RFM_ERR_INVALID;
  std::lock_guard<std::mutex> lock(m->mu);
  if (!m->st_valid)
    return RFM_ERR_INVALID; // the reference returns false until the first DemuxRead
  // the reference's operands are float: log10 resolves to the float overload, the product by 20 is a float product, and
  // only "+ 3.01" is done in double (pinned against the compiled add-on: tests/test_ref_addon.py, test_gpu_demux.py)
  if (interface_level)
    *interface_level = 20 * log10f(m->st_interface_level);
  if (audio_level_db)
    *audio_level_db = (float)(20 * log10f(m->st_audio_level) + 3.01);
  if (stereo)
    *stereo = m->st_stereo;
  return RFM_OK;
}

} // extern "C"
