// dp.cpp — one host process driving several B200s: the data-parallel mode of include/mnv1.h.
//
// The reference runs ONE image on ONE in-order queue (MobileNet.c:29,171-173); images are independent, so
// BASELINE config 5 cuts the batch into contiguous shards.  A group owns one mnv1_ctx and one
// worker thread per GPU; an entry point posts the same call to every worker and joins.  Nothing is
// exchanged during the forward pass; the logits gather is done by each rank's head kernel storing its
// rows into every rank's gather block over NVLink (mnv1_gather_*, head.cu), so there is no collective
// kernel in the step and nothing here links NCCL.  Built on the public C-ABI only.
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mnv1.h"

namespace {

class Worker {   // a thread that runs posted jobs in order
 public:
  Worker() : th_([this] { loop(); }) {}
  ~Worker() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
    cv_.notify_all();
    th_.join();
  }
  void post(std::function<int()> job) {
    { std::lock_guard<std::mutex> lk(mu_); job_ = std::move(job); has_job_ = true; done_ = false; }
    cv_.notify_all();
  }
  int join() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [this] { return done_; });
    return rc_;
  }

 private:
  void loop() {
    for (;;) {
      std::function<int()> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return has_job_ || stop_; });
        if (stop_) return;
        job = std::move(job_);
        has_job_ = false;
      }
      const int rc = job();
      { std::lock_guard<std::mutex> lk(mu_); rc_ = rc; done_ = true; }
      cv_.notify_all();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::function<int()> job_;
  bool has_job_ = false, done_ = true, stop_ = false;
  int rc_ = 0;
  std::thread th_;
};

}  // namespace

struct mnv1_dp {
  int n = 0;
  int rows_per_gpu = 0;
  std::vector<mnv1_ctx*> ctx;
  std::vector<std::unique_ptr<Worker>> workers;
  std::string err;
  std::vector<float> weight;   // shard weights (mnv1_dp_set_shard_weights / mnv1_dp_calibrate); empty = equal shards
  // tickets of mnv1_dp_forward_submit: per-rank tickets of the last kRing submits
  static constexpr int kRing = 8;
  long next_ticket = 0;
  std::vector<long> rank_ticket[kRing];
  long ring_ticket[kRing];

  // run fn(rank) on every worker, return the first failure (and remember its message)
  int all(const std::function<int(int)>& fn) {
    for (int r = 0; r < n; ++r) workers[r]->post([&fn, r] { return fn(r); });
    int rc = MNV1_OK;
    for (int r = 0; r < n; ++r) {
      const int c = workers[r]->join();
      if (c != MNV1_OK && rc == MNV1_OK) {
        rc = c;
        err = "rank " + std::to_string(r) + ": " + mnv1_last_error(ctx[r]);
      }
    }
    return rc;
  }
};

static thread_local std::string g_dp_error;

extern "C" {

const char* mnv1_dp_last_error(const mnv1_dp* dp) { return dp ? dp->err.c_str() : g_dp_error.c_str(); }
int mnv1_dp_size(const mnv1_dp* dp) { return dp ? dp->n : 0; }
mnv1_ctx* mnv1_dp_ctx(mnv1_dp* dp, int rank) { return (dp && rank >= 0 && rank < dp->n) ? dp->ctx[rank] : nullptr; }

int mnv1_dp_shard(int n, int rank, int world, int* first, int* count) {
  if (world <= 0 || rank < 0 || rank >= world || n < 0 || !first || !count) return MNV1_EINVAL;
  const int base = n / world, rem = n % world;
  *first = rank * base + (rank < rem ? rank : rem);
  *count = base + (rank < rem ? 1 : 0);
  return MNV1_OK;
}

int mnv1_dp_shard_weighted(int n, int rank, int world, const float* weight, int* first, int* count) {
  if (!weight) return mnv1_dp_shard(n, rank, world, first, count);
  if (world <= 0 || world > 64 || rank < 0 || rank >= world || n < 0 || !first || !count) return MNV1_EINVAL;
  double sum = 0;
  for (int r = 0; r < world; ++r) {
    if (!(weight[r] > 0.f)) return MNV1_EINVAL;   // also rejects NaN
    sum += weight[r];
  }
  // largest-remainder apportionment: floor of the exact share, the leftover images to the largest fractions
  // (ties to the lower rank), so the counts add up to n and differ from the exact share by less than one image
  int cnt[64];
  double frac[64];
  int given = 0;
  for (int r = 0; r < world; ++r) {
    const double share = (double)n * weight[r] / sum;
    cnt[r] = (int)share;
    frac[r] = share - cnt[r];
    given += cnt[r];
  }
  for (; given < n; ++given) {
    int best = 0;
    for (int r = 1; r < world; ++r)
      if (frac[r] > frac[best]) best = r;
    ++cnt[best];
    frac[best] = -1.0;
  }
  int f = 0;
  for (int r = 0; r < rank; ++r) f += cnt[r];
  *first = f;
  *count = cnt[rank];
  return MNV1_OK;
}

int mnv1_dp_destroy(mnv1_dp* dp) {
  if (!dp) return MNV1_OK;
  for (int r = 0; r < (int)dp->ctx.size(); ++r)
    if (dp->ctx[r]) mnv1_sync(dp->ctx[r]);
  dp->workers.clear();                      // joins the threads
  for (auto* c : dp->ctx) mnv1_ctx_destroy(c);
  delete dp;
  return MNV1_OK;
}

int mnv1_dp_create(const int* devices, int n_devices, mnv1_dtype dtype, int max_batch_per_gpu, mnv1_dp** out) {
  if (!out) return MNV1_EINVAL;
  *out = nullptr;
  if (!devices || n_devices < 1 || n_devices > 8 || max_batch_per_gpu <= 0) {
    g_dp_error = "dp_create: need 1..8 devices and a positive per-GPU batch";
    return MNV1_EINVAL;
  }
  std::unique_ptr<mnv1_dp> dp(new mnv1_dp);
  dp->n = n_devices;
  dp->rows_per_gpu = max_batch_per_gpu;
  dp->ctx.assign(n_devices, nullptr);
  for (int i = 0; i < mnv1_dp::kRing; ++i) { dp->ring_ticket[i] = -1; dp->rank_ticket[i].assign(n_devices, -1); }
  auto bail = [&](int rc, const std::string& msg) {
    g_dp_error = msg;
    mnv1_dp_destroy(dp.release());
    return rc;
  };
  for (int r = 0; r < n_devices; ++r) {
    int rc = mnv1_ctx_create(devices[r], dtype, &dp->ctx[r]);
    if (rc) return bail(rc, std::string("dp_create: device ") + std::to_string(devices[r]) + ": " + mnv1_last_error(nullptr));
    dp->workers.emplace_back(new Worker);
  }
  if (dtype != MNV1_U8) {   // the gather stores live in the softmax kernel of the fp32 / bf16 head
    for (int r = 0; r < n_devices; ++r) {
      int rc = mnv1_gather_create(dp->ctx[r], n_devices, r, max_batch_per_gpu);
      if (rc) return bail(rc, std::string("dp_create: ") + mnv1_last_error(dp->ctx[r]));
    }
    for (int r = 0; r < n_devices; ++r)
      for (int d = 0; d < n_devices; ++d) {
        if (d == r) continue;
        int rc = mnv1_gather_attach(dp->ctx[r], dp->ctx[d]);
        if (rc) return bail(rc, std::string("dp_create: ") + mnv1_last_error(dp->ctx[r]));
      }
  }
  *out = dp.release();
  return MNV1_OK;
}

int mnv1_dp_set_pad_mode(mnv1_dp* dp, mnv1_pad pad) {
  if (!dp) return MNV1_EINVAL;
  return dp->all([&](int r) { return mnv1_ctx_set_pad_mode(dp->ctx[r], pad); });
}
int mnv1_dp_set_input_transform(mnv1_dp* dp, float scale, float bias) {
  if (!dp) return MNV1_EINVAL;
  return dp->all([&](int r) { return mnv1_ctx_set_input_transform(dp->ctx[r], scale, bias); });
}
int mnv1_dp_set_weights(mnv1_dp* dp, const float* weights, const float* scale, const float* shift, mnv1_act act) {
  if (!dp) return MNV1_EINVAL;
  int rc = dp->all([&](int r) { return mnv1_set_weights(dp->ctx[r], weights, scale, shift, act); });   // replicated
  if (rc) return rc;
  return dp->all([&](int r) { return mnv1_plan(dp->ctx[r], dp->rows_per_gpu); });
}
int mnv1_dp_load_weights(mnv1_dp* dp, const char* path, mnv1_act act) {
  if (!dp || !path) return MNV1_EINVAL;
  std::vector<float> w(MNV1_TOTAL_WEIGHTS), sc(MNV1_BN_CHANNELS + 1000), sh(MNV1_BN_CHANNELS + 1000);
  int rc = mnv1_parse_weights(path, w.data(), sc.data(), sh.data());   // the file is read once, not once per GPU
  if (rc) { dp->err = mnv1_last_error(nullptr); return rc; }
  return mnv1_dp_set_weights(dp, w.data(), sc.data(), sh.data(), act);
}

static const size_t kImg = 224 * 224 * 3;

int mnv1_dp_set_shard_weights(mnv1_dp* dp, const float* weight) {
  if (!dp) return MNV1_EINVAL;
  if (!weight) { dp->weight.clear(); return MNV1_OK; }
  for (int r = 0; r < dp->n; ++r)
    if (!(weight[r] > 0.f)) { dp->err = "dp_set_shard_weights: weights must be positive"; return MNV1_EINVAL; }
  dp->weight.assign(weight, weight + dp->n);
  return MNV1_OK;
}

// The host feeds its GPUs unevenly when they all copy at once (one memory controller nearer than the other, shared
// PCIe switches; a guest may not even see the topology): measure every GPU's pinned H2D rate with all of them copying
// and cut the batch in proportion, so that the uploads of a step finish together.
int mnv1_dp_calibrate(mnv1_dp* dp, float* gbytes_per_s) {
  if (!dp) return MNV1_EINVAL;
  std::vector<float> rate(dp->n, 0.f);
  std::vector<mnv1_h2d_probe_t*> probe(dp->n, nullptr);
  const size_t bytes = (size_t)dp->rows_per_gpu * kImg;
  // allocate everywhere first: pinning takes tens of milliseconds and the copies must overlap
  int rc = dp->all([&](int r) { return mnv1_h2d_probe_open(dp->ctx[r], bytes, &probe[r]); });
  if (!rc) rc = dp->all([&](int r) { return mnv1_h2d_probe_run(probe[r], 12, &rate[r]); });
  if (!rc) {
    // the GPUs with the fast links finished early and left the others alone with the memory system: repeat with
    // copy counts in proportion to the first rates, so that everybody copies until the end
    float top = 0.f;
    for (float x : rate) top = x > top ? x : top;
    std::vector<int> reps(dp->n);
    for (int r = 0; r < dp->n; ++r) reps[r] = (int)(30.f * rate[r] / top + 0.5f) < 4 ? 4 : (int)(30.f * rate[r] / top + 0.5f);
    rc = dp->all([&](int r) { return mnv1_h2d_probe_run(probe[r], reps[r], &rate[r]); });
  }
  dp->all([&](int r) { return mnv1_h2d_probe_close(probe[r]); });
  if (rc) return rc;
  if (gbytes_per_s) std::memcpy(gbytes_per_s, rate.data(), sizeof(float) * dp->n);
  return mnv1_dp_set_shard_weights(dp, rate.data());
}

int mnv1_dp_forward_submit(mnv1_dp* dp, const uint8_t* images, int n, float* logits, int* top1, float* top1_prob,
                           long* ticket) {
  if (!dp || !images || !ticket || n <= 0) return MNV1_EINVAL;
  const float* wt = dp->weight.empty() ? nullptr : dp->weight.data();
  for (int r = 0; r < dp->n; ++r) {
    int first = 0, count = 0;
    mnv1_dp_shard_weighted(n, r, dp->n, wt, &first, &count);
    if (count > dp->rows_per_gpu) { dp->err = "dp_forward: a shard exceeds max_batch_per_gpu"; return MNV1_EINVAL; }
  }
  const int slot = (int)(dp->next_ticket % mnv1_dp::kRing);
  std::vector<long>& rt = dp->rank_ticket[slot];
  int rc = dp->all([&](int r) {
    int first = 0, count = 0;
    mnv1_dp_shard_weighted(n, r, dp->n, wt, &first, &count);
    rt[r] = -1;
    if (count == 0) return MNV1_OK;
    return mnv1_forward_submit(dp->ctx[r], images + (size_t)first * kImg, count,
                               logits ? logits + (size_t)first * MNV1_NUM_CLASSES : nullptr, top1 ? top1 + first : nullptr,
                               top1_prob ? top1_prob + first : nullptr, &rt[r]);
  });
  if (rc) return rc;
  dp->ring_ticket[slot] = dp->next_ticket;
  *ticket = dp->next_ticket++;
  return MNV1_OK;
}

int mnv1_dp_forward_wait(mnv1_dp* dp, long ticket) {
  if (!dp || ticket < 0 || ticket >= dp->next_ticket) return MNV1_EINVAL;
  const int slot = (int)(ticket % mnv1_dp::kRing);
  if (dp->ring_ticket[slot] != ticket) return MNV1_OK;   // older than the ring: retired by later submits of every rank
  const std::vector<long>& rt = dp->rank_ticket[slot];
  return dp->all([&](int r) { return rt[r] < 0 ? MNV1_OK : mnv1_forward_wait(dp->ctx[r], rt[r]); });
}

int mnv1_dp_forward(mnv1_dp* dp, const uint8_t* images, int n, float* logits, int* top1, float* top1_prob) {
  long t = -1;
  int rc = mnv1_dp_forward_submit(dp, images, n, logits, top1, top1_prob, &t);
  if (rc) return rc;
  return mnv1_dp_forward_wait(dp, t);
}

int mnv1_dp_forward_device(mnv1_dp* dp, const void* const* d_images, int n_per_gpu) {
  if (!dp || !d_images || n_per_gpu <= 0 || n_per_gpu > dp->rows_per_gpu) return MNV1_EINVAL;
  int rc = dp->all([&](int r) {
    void *lg = nullptr, *t1 = nullptr, *p1 = nullptr;
    int c = mnv1_gather_ptrs(dp->ctx[r], &lg, &t1, &p1);
    if (c) return c;
    // a rank's own rows of its gather block double as its local outputs
    const size_t row0 = (size_t)r * dp->rows_per_gpu;
    return mnv1_forward_device(dp->ctx[r], d_images[r], n_per_gpu, (float*)lg + row0 * MNV1_NUM_CLASSES, (int*)t1 + row0,
                               (float*)p1 + row0);
  });
  if (rc) return rc;
  return dp->all([&](int r) { return mnv1_sync(dp->ctx[r]); });   // every rank's peer stores have landed
}

}  // extern "C"
