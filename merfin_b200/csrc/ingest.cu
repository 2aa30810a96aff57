// Hybrid ingest for the host-buffer -hist path: two lanes feed the packed sequence arrays.
//
//   lane G  the piece's ASCII goes over PCIe as is (1 byte per base) and is packed by k_pack
//   lane C  host threads pack the piece with the SIMD packer (host_pack.cpp) into pinned
//           staging, and only the packed words (0.375 byte per base) cross PCIe
//
// Pieces are tile-aligned ranges of global positions, so the two lanes never share a word.
// The share of pieces that goes to lane C follows the rates measured in the previous call.
#include "table_host.h"

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

extern "C" void mfn_pack_host_range(const char *bases, const uint64_t *rec_start, uint32_t nseq, uint64_t g_begin,
                                    uint64_t g_end, uint64_t *seq2_out, uint64_t *ok_out);

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct IngestPiece {
  uint64_t g_begin = 0, g_end = 0;
  int lane = 0;                       // 0 = G, 1 = C
  cudaEvent_t done = nullptr;         // recorded when the piece's packed words are (being) written
  // lane C
  int buf = 0, prev = -1;             // staging buffer, and the earlier piece that used it last
  uint32_t tasks_left = 0;
  bool buf_ready = false, buf_checking = false, uploaded = false;
};

struct mfn_ingest {
  mfn_ctx *ctx = nullptr;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  bool stop = false;
  bool abort = false;                                     // the current call is being abandoned (error path)
  // current call
  mfn_seq *seq = nullptr;
  const char *bases = nullptr;
  std::vector<IngestPiece> pieces;
  std::deque<std::pair<uint32_t, uint32_t>> tasks;       // (piece, sub-task)
  uint64_t sub = 1 << 20;                                 // positions per sub-task (multiple of 64)
  int error = 0;
  uint32_t active = 0;                                    // workers inside a task
  double t0 = 0, t_last_c = 0;
  double busy_c = 0;                                      // seconds the workers spent packing (summed over threads)
  uint64_t bytes_c = 0, bytes_g = 0;
  uint32_t fed = 0, last_g = 0;                           // lane G pieces queued so far
  bool stamped = false;
  // pinned staging ring for packed pieces
  static constexpr int NBUF = 6;
  unsigned long long *h_stage[NBUF] = {};
  size_t stage_positions = 0;                             // capacity of one buffer in positions
  cudaStream_t up_stream = nullptr;                       // packed uploads
  std::vector<cudaEvent_t> events;
  size_t events_used = 0;
  // lane balance: share of positions packed on the host, and the timestamp of the last G piece
  double cpu_share = -1.0;
  double t_last_g = 0;
};

static void ingest_worker(mfn_ingest *I) {
  cudaSetDevice(I->ctx->device);
  std::unique_lock<std::mutex> lk(I->mu);
  while (true) {
    I->cv_work.wait(lk, [&] { return I->stop || !I->tasks.empty(); });
    if (I->stop) return;
    const auto task = I->tasks.front();
    I->tasks.pop_front();
    IngestPiece &P = I->pieces[task.first];
    I->active++;
    // the staging buffer is free once the upload of the piece that used it last has completed
    while (!P.buf_ready && !I->stop && !I->abort) {
      if (P.prev < 0) { P.buf_ready = true; break; }
      if (!P.buf_checking) {
        P.buf_checking = true;
        IngestPiece &Q = I->pieces[P.prev];
        I->cv_done.wait(lk, [&] { return Q.uploaded || I->stop || I->abort || I->error; });
        if (I->stop || I->abort) break;
        cudaEvent_t ev = Q.done;
        lk.unlock();
        cudaEventSynchronize(ev);
        lk.lock();
        P.buf_ready = true;
        I->cv_done.notify_all();
      } else {
        I->cv_done.wait(lk, [&] { return P.buf_ready || I->stop || I->abort; });
      }
    }
    if (I->stop) return;
    if (I->abort) { I->active--; I->cv_done.notify_all(); continue; }
    lk.unlock();
    const uint64_t a = P.g_begin + (uint64_t)task.second * I->sub, b = std::min(P.g_end, a + I->sub);
    unsigned long long *s2 = I->h_stage[P.buf], *so = s2 + I->stage_positions / 32;
    const double tp = now_s();
    mfn_pack_host_range(I->bases, I->seq->rec_start.data(), I->seq->nseq, a, b,
                        (uint64_t *)s2 + (a - P.g_begin) / 32, (uint64_t *)so + (a - P.g_begin) / 64);
    const double dtp = now_s() - tp;
    lk.lock();
    I->busy_c += dtp;
    if (--P.tasks_left == 0) {
      // last sub-task of the piece: upload the packed words straight into the device arrays
      lk.unlock();
      const uint64_t n = P.g_end - P.g_begin;
      cudaError_t e = cudaMemcpyAsync(I->seq->d_seq2 + P.g_begin / 32, s2, n / 32 * 8, cudaMemcpyHostToDevice, I->up_stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(I->seq->d_ok + P.g_begin / 64, so, n / 64 * 8, cudaMemcpyHostToDevice, I->up_stream);
      if (e == cudaSuccess) e = cudaEventRecord(P.done, I->up_stream);
      I->ctx->h2d_bytes += n / 32 * 8 + n / 64 * 8;
      lk.lock();
      if (e != cudaSuccess) I->error = 1;
      P.uploaded = true;
      I->t_last_c = now_s();
    }
    I->active--;
    I->cv_done.notify_all();
  }
}

static void CUDART_CB ingest_stamp(void *p) { *(double *)p = now_s(); }

void mfn_ingest_free(mfn_ingest *I) {
  if (!I) return;
  {
    std::lock_guard<std::mutex> g(I->mu);
    I->stop = true;
  }
  I->cv_work.notify_all();
  I->cv_done.notify_all();
  for (auto &t : I->workers) t.join();
  for (auto &h : I->h_stage) if (h) cudaFreeHost(h);
  for (auto &e : I->events) cudaEventDestroy(e);
  if (I->up_stream) cudaStreamDestroy(I->up_stream);
  delete I;
}

static int ingest_get(mfn_ctx *ctx, mfn_ingest **out) {
  if (!ctx->ingest) {
    mfn_ingest *I = new mfn_ingest();
    I->ctx = ctx;
    const char *env = getenv("MFN_HOST_PACK_THREADS");
    int hw = (int)std::thread::hardware_concurrency();
    int nt = env ? atoi(env) : std::max(0, std::min(hw - 2, 30));
    if (nt < 0) nt = 0;
    MFN_CUDA(cudaStreamCreateWithFlags(&I->up_stream, cudaStreamNonBlocking));
    for (int i = 0; i < nt; i++) I->workers.emplace_back(ingest_worker, I);
    ctx->ingest = I;
  }
  *out = ctx->ingest;
  return MFN_OK;
}

static cudaEvent_t ingest_event(mfn_ingest *I) {
  if (I->events_used == I->events.size()) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    I->events.push_back(e);
  }
  return I->events[I->events_used++];
}

// Start packing tiles [tile_begin, tile_end) of `seq` (freshly zeroed arrays) cut into pieces of
// `piece_tiles` tiles (the last piece also covers one tile of halo).  Everything lane G needs is
// queued here; lane C runs on the worker threads.  mfn_ingest_wait(i) then makes the compute
// stream wait for piece i.
int mfn_ingest_begin(mfn_ctx *ctx, mfn_seq *seq, const char *bases, uint64_t tile_begin, uint64_t tile_end,
                     uint64_t piece_tiles, uint32_t *npieces) {
  mfn_ingest *I = nullptr;
  MFN_TRY(ingest_get(ctx, &I));
  const uint64_t ntiles = tile_end - tile_begin;
  const uint32_t np = (uint32_t)std::max<uint64_t>(1, ceil_div_u64(ntiles, piece_tiles));
  const int nt = (int)I->workers.size();
  // share of lane C: measured rates of the previous call; a model of the two lanes the first time
  const char *env = getenv("MFN_HOST_PACK_SHARE");
  double share = env ? atof(env) : I->cpu_share;
  if (share < 0) share = nt ? (3.0 * nt) / (3.0 * nt + 50.0) : 0.0;
  if (nt == 0 || np < 4) share = 0.0;
  share = std::min(share, 0.9);
  const size_t piece_pos = (size_t)(piece_tiles + 1) * MFN_TILE;
  if (share > 0 && I->stage_positions < piece_pos) {
    for (auto &h : I->h_stage) { if (h) cudaFreeHost(h); h = nullptr; }
    for (auto &h : I->h_stage) MFN_CUDA(cudaMallocHost(&h, piece_pos / 32 * 8 + piece_pos / 64 * 8));
    I->stage_positions = piece_pos;
  }
  {
    std::lock_guard<std::mutex> g(I->mu);
    I->seq = seq; I->bases = bases; I->error = 0; I->events_used = 0;
    I->pieces.assign(np, IngestPiece());
    I->tasks.clear();
    I->t_last_c = 0; I->t_last_g = 0; I->stamped = false; I->fed = 0; I->busy_c = 0;
    double acc = 0.0;
    int nextbuf = 0;
    for (uint32_t i = 0; i < np; i++) {
      IngestPiece &P = I->pieces[i];
      P.g_begin = (tile_begin + (uint64_t)i * piece_tiles) * MFN_TILE;
      P.g_end = (i + 1 == np ? std::min(tile_end + 1, seq->ntiles + 1) : tile_begin + (uint64_t)(i + 1) * piece_tiles) * MFN_TILE;
      acc += share;
      P.lane = acc >= 1.0 ? 1 : 0;
      if (P.lane) {
        acc -= 1.0;
        P.buf = nextbuf; nextbuf = (nextbuf + 1) % mfn_ingest::NBUF;
        for (int q = (int)i - 1; q >= 0; q--) if (I->pieces[q].lane && I->pieces[q].buf == P.buf) { P.prev = q; break; }
      }
      P.done = ingest_event(I);
      if (!P.done) return MFN_ERR_CUDA;
    }
  }
  // the packed arrays were zeroed on the pack stream: uploads of packed words come after that
  MFN_CUDA(cudaEventRecord(ctx->ev[MFN_STAGE_BUFS + 2], ctx->pack_stream));
  MFN_CUDA(cudaStreamWaitEvent(I->up_stream, ctx->ev[MFN_STAGE_BUFS + 2], 0));
  I->t0 = now_s();
  // lane C: hand the sub-tasks to the workers, in piece order
  uint64_t bytes_c = 0, bytes_g = 0;
  {
    std::lock_guard<std::mutex> g(I->mu);
    for (uint32_t i = 0; i < np; i++) {
      IngestPiece &P = I->pieces[i];
      if (!P.lane) { bytes_g += P.g_end - P.g_begin; continue; }
      bytes_c += P.g_end - P.g_begin;
      const uint32_t n = (uint32_t)ceil_div_u64(P.g_end - P.g_begin, I->sub);
      P.tasks_left = n;
      for (uint32_t s = 0; s < n; s++) I->tasks.emplace_back(i, s);
    }
  }
  I->bytes_c = bytes_c; I->bytes_g = bytes_g;
  I->cv_work.notify_all();
  I->fed = 0;
  *npieces = np;
  return MFN_OK;
}

// Lane G: queue the raw upload + k_pack of the pieces below `upto` that are not queued yet.  The
// H2D engine serves copies in submission order, so lane G is fed a few pieces ahead of the scan
// only: packed uploads of lane C must not wait behind the whole assembly.
int mfn_ingest_feed(mfn_ctx *ctx, uint32_t upto) {
  mfn_ingest *I = ctx->ingest;
  const uint32_t np = (uint32_t)I->pieces.size();
  for (; I->fed < std::min(upto, np); I->fed++) {
    IngestPiece &P = I->pieces[I->fed];
    if (P.lane) continue;
    MFN_TRY(mfn_seq_pack_range(I->seq, I->bases, P.g_begin, P.g_end, P.done));
    I->last_g = I->fed;
  }
  if (I->fed == np && I->bytes_g && !I->stamped) {
    I->stamped = true;
    MFN_CUDA(cudaLaunchHostFunc(ctx->pack_stream, ingest_stamp, &I->t_last_g));
  }
  return MFN_OK;
}

// Make the compute stream wait until piece i is packed on the device.
int mfn_ingest_wait(mfn_ctx *ctx, uint32_t i) {
  mfn_ingest *I = ctx->ingest;
  IngestPiece &P = I->pieces[i];
  if (!P.lane && I->fed <= i) MFN_TRY(mfn_ingest_feed(ctx, i + 1));
  if (P.lane) {
    std::unique_lock<std::mutex> lk(I->mu);
    I->cv_done.wait(lk, [&] { return P.uploaded || I->error; });
    if (I->error) { mfn_set_error("host packing lane: upload failed"); return MFN_ERR_CUDA; }
  }
  MFN_CUDA(cudaStreamWaitEvent(ctx->stream, P.done, 0));
  return MFN_OK;
}

// After the last piece: rebalance the lanes from what this call measured.
void mfn_ingest_end(mfn_ctx *ctx) {
  mfn_ingest *I = ctx->ingest;
  if (!I) return;
  {
    // nothing of this call may still run on the workers (also the way out after an error)
    std::unique_lock<std::mutex> lk(I->mu);
    const bool unfinished = !I->tasks.empty();
    I->tasks.clear();
    if (unfinished) { I->abort = true; I->cv_done.notify_all(); }
    I->cv_done.wait(lk, [&] { return I->active == 0; });
    I->abort = false;
  }
  if (!I->bytes_c || I->busy_c <= 0 || I->workers.empty()) return;
  // Balance the lanes.  Host packing rate rc = bytes / (busy time / threads); the link carries
  // 1 byte per base of lane G and 0.375 of lane C at rp bytes/s.  Both lanes finish together when
  // f / rc = (1 - 0.625 f) / rp, i.e. f = rc / (rp + 0.625 rc).
  const char *env = getenv("MFN_PCIE_GBS");
  const double rp = (env ? atof(env) : 52.0) * 1e9;
  const double rc = (double)I->bytes_c / (I->busy_c / (double)I->workers.size());
  const double target = rc / (rp + 0.625 * rc);
  const double cur = (double)I->bytes_c / (double)(I->bytes_c + I->bytes_g);
  I->cpu_share = std::min(0.9, std::max(0.05, 0.5 * cur + 0.5 * target));
  if (getenv("MFN_INGEST_DEBUG")) {
    cudaStreamSynchronize(ctx->pack_stream);               // t_last_g is stamped by a host function on it
    fprintf(stderr, "[ingest] lane G %.2f GB done at %.1f ms, lane C %.2f GB done at %.1f ms, host packing %.1f GB/s on %zu threads -> host share %.3f\n",
            I->bytes_g / 1e9, (I->t_last_g - I->t0) * 1e3, I->bytes_c / 1e9, (I->t_last_c - I->t0) * 1e3, rc / 1e9, I->workers.size(), I->cpu_share);
  }
}
