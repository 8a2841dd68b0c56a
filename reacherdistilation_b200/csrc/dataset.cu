// Device-resident rollout buffer with the reference Dataset's semantics (/root/reference src/distilation/dataset.py:72-296):
//   write()            dataset.py:118-143   append one record {ob, rew, t, s, with} per env (prev / prew are derived, see below)
//   flush()            dataset.py:146-149   close the current episodes
//   training_batches() dataset.py:179-210   B episodes drawn with replacement + ONE shared random start, T-step windows, time-major
//   test_batch()       dataset.py:213-290   tail window of the current episode + the current observation
// The reference keeps Python lists of dicts and pages them to gzip JSON; here N lock-step envs append one record each per step
// into a ring of G "generations" (one generation = the N episodes that run concurrently), laid out
//     field[(slot * 50 + k) * N + env]           slot = generation % G, k = step in episode
// so a write is a contiguous copy and window gathers read contiguous rows.  The derived fields of a record j are
//     prev[j] = t[j-1], prew[j] = rew[j-1], zeros for j == 0      (dataset.py:128-129,151-164 pdflat_at / rew_at)
// Sampling uses Philox4x32-10 keyed (seed; draw, b, 0, STREAM_DATASET) with multiply-shift range reduction -- same code in
// oracle/dataset_np.py -- instead of Python's `random` (unpinned in the reference).
#include "common.cuh"
#include "philox.cuh"

struct rb_dataset {
    int64_t n = 0;          // envs
    int64_t G = 0;          // generations kept
    int device = 0;
    int k = 0;              // records written in the current (open) generation
    int64_t gen = 0;        // generations flushed so far
    float* ob = nullptr;    // [G*50*N, 11]
    float* rew = nullptr;   // [G*50*N]
    float4* t = nullptr;    // [G*50*N]
    float4* s = nullptr;    // [G*50*N]
    uint8_t* with = nullptr;// [G*50*N]   0 = 't', 1 = 's'
};

namespace rb {

constexpr uint32_t STREAM_DATASET = 3u;
constexpr int EP = RB_EPISODE_STEPS;

__global__ void k_ds_write(int64_t n, int64_t row0, const float* __restrict__ ob, const float* __restrict__ rew, const float4* __restrict__ t,
                           const float4* __restrict__ s, uint8_t with, float* __restrict__ d_ob, float* __restrict__ d_rew, float4* __restrict__ d_t,
                           float4* __restrict__ d_s, uint8_t* __restrict__ d_with) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * 11) d_ob[row0 * 11 + i] = __ldg(ob + i);
    if (i < n) {
        d_rew[row0 + i] = rew ? __ldg(rew + i) : 0.f;
        d_t[row0 + i] = t ? __ldg(t + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        d_s[row0 + i] = s ? __ldg(s + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        d_with[row0 + i] = with;
    }
}

// episode id (0 .. avail*N-1, oldest generation first) -> (slot, env)
__device__ __forceinline__ void episode_slot(int64_t ep, int64_t n, int64_t G, int64_t gen, int64_t avail, int64_t& slot, int64_t& env) {
    const int64_t g = ep / n;
    env = ep - g * n;
    slot = (gen - avail + g) % G;
}

__global__ void k_ds_training_batch(const rb_dataset d, int64_t avail, uint32_t k0, uint32_t k1, uint32_t draw, int B, int T, float* __restrict__ ob_out,
                                    float4* __restrict__ t_out, float4* __restrict__ prev_out, float* __restrict__ prew_out,
                                    int32_t* __restrict__ episodes_out, int32_t* __restrict__ start_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // idx = tau * B + b
    if (idx >= T * B) return;
    const int tau = idx / B, b = idx - tau * B;
    const uint32_t xs = philox4x32_10(draw, 0xFFFFFFFFu, 0u, STREAM_DATASET, k0, k1).x;                   // the shared start (dataset.py:187)
    const int start = (int)(((uint64_t)xs * (uint64_t)(EP - T + 1)) >> 32);
    const uint32_t xe = philox4x32_10(draw, (uint32_t)b, 0u, STREAM_DATASET, k0, k1).x;                   // random.choice (dataset.py:186)
    const int64_t ep = (int64_t)(((uint64_t)xe * (uint64_t)(avail * d.n)) >> 32);
    int64_t slot, env;
    episode_slot(ep, d.n, d.G, d.gen, avail, slot, env);
    const int j = start + tau;
    const int64_t r = (slot * EP + j) * d.n + env;
#pragma unroll
    for (int c = 0; c < 11; ++c) ob_out[(int64_t)idx * 11 + c] = d.ob[r * 11 + c];
    t_out[idx] = d.t[r];
    prev_out[idx] = j > 0 ? d.t[r - d.n] : make_float4(0.f, 0.f, 0.f, 0.f);
    prew_out[idx] = j > 0 ? d.rew[r - d.n] : 0.f;
    if (tau == 0) {
        if (episodes_out) episodes_out[b] = (int32_t)ep;
        if (start_out && b == 0) *start_out = start;
    }
}

// test batch for every env (the batched form of dataset.py:213-290): column e = tail window of env e's current episode
//   ob[tau]   = ob_rec[len - T + 1 + tau] (tau < T-1; zeros where the index is negative), ob[T-1] = current observation
//   prev[tau] = t_rec[len - T + tau],  prew[tau] = rew_rec[len - T + tau]   (zeros where negative)
__global__ void k_ds_test_batch(const rb_dataset d, int64_t slot, int len, int T, const float* __restrict__ ob_cur, float* __restrict__ ob_out,
                                float4* __restrict__ prev_out, float* __restrict__ prew_out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;    // idx = tau * N + e
    if (idx >= (int64_t)T * d.n) return;
    const int tau = (int)(idx / d.n);
    const int64_t e = idx - (int64_t)tau * d.n;
    const int jo = len - T + 1 + tau, jp = len - T + tau;
    if (tau == T - 1) {
#pragma unroll
        for (int c = 0; c < 11; ++c) ob_out[idx * 11 + c] = __ldg(ob_cur + e * 11 + c);
    } else {
        const int64_t r = (slot * EP + jo) * d.n + e;
#pragma unroll
        for (int c = 0; c < 11; ++c) ob_out[idx * 11 + c] = jo >= 0 ? d.ob[r * 11 + c] : 0.f;
    }
    const int64_t rp = (slot * EP + jp) * d.n + e;
    if (prev_out) prev_out[idx] = jp >= 0 ? d.t[rp] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (prew_out) prew_out[idx] = jp >= 0 ? d.rew[rp] : 0.f;
}

}  // namespace rb

using namespace rb;

extern "C" {

int rb_dataset_destroy(rb_dataset* d) {
    if (!d) return RB_OK;
    DeviceGuard guard(d->device);
    cudaFree(d->ob); cudaFree(d->rew); cudaFree(d->t); cudaFree(d->s); cudaFree(d->with);
    delete d;
    return RB_OK;
}

int rb_dataset_create(rb_dataset** out, int64_t num_envs, int64_t generations, int device) {
    RB_REQUIRE(out != nullptr, "out is NULL");
    RB_REQUIRE(num_envs > 0 && generations > 0, "num_envs and generations must be positive");
    DeviceGuard guard(device);
    rb_dataset* d = new rb_dataset();
    d->n = num_envs; d->G = generations; d->device = device;
    const size_t rows = (size_t)generations * EP * num_envs;
    cudaError_t err = cudaMalloc(&d->ob, rows * 11 * sizeof(float));
    if (err == cudaSuccess) err = cudaMalloc(&d->rew, rows * sizeof(float));
    if (err == cudaSuccess) err = cudaMalloc(&d->t, rows * sizeof(float4));
    if (err == cudaSuccess) err = cudaMalloc(&d->s, rows * sizeof(float4));
    if (err == cudaSuccess) err = cudaMalloc(&d->with, rows);
    if (err != cudaSuccess) { rb_dataset_destroy(d); return cuda_fail(err, "rb_dataset_create"); }
    *out = d;
    return RB_OK;
}

int rb_dataset_write(rb_dataset* d, const float* ob, const float* rew, const float* t, const float* s, int stepped_with, void* stream) {
    RB_REQUIRE(d && ob, "NULL argument");
    RB_REQUIRE(d->k < EP, "the current episodes already hold EPISODE_STEPS records: flush() first");
    RB_REQUIRE(stepped_with == 0 || stepped_with == 1, "stepped_with: 0 = 't', 1 = 's'");
    const int64_t row0 = ((d->gen % d->G) * EP + d->k) * d->n;
    k_ds_write<<<(unsigned)((d->n * 11 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d->n, row0, ob, rew, (const float4*)t, (const float4*)s,
                                                                                     (uint8_t)stepped_with, d->ob, d->rew, d->t, d->s, d->with);
    RB_CUDA(cudaGetLastError());
    d->k += 1;
    return RB_OK;
}

int rb_dataset_flush(rb_dataset* d) {
    RB_REQUIRE(d != nullptr, "NULL argument");
    RB_REQUIRE(d->k == EP, "flush() closes complete episodes only (EPISODE_STEPS records; the reference flushes on done)");
    d->gen += 1; d->k = 0;
    return RB_OK;
}

int64_t rb_dataset_num_episodes(const rb_dataset* d) { return d ? d->gen * d->n : 0; }
int64_t rb_dataset_num_available(const rb_dataset* d) { return d ? (d->gen < d->G ? d->gen : d->G - (d->k > 0 ? 1 : 0)) * d->n : 0; }
int rb_dataset_episode_len(const rb_dataset* d) { return d ? d->k : 0; }

int64_t rb_dataset_generations(const rb_dataset* d) { return d ? d->gen : 0; }

int rb_dataset_export_host(rb_dataset* d, int64_t generation, float* ob_host, float* rew_host, float* t_host, float* s_host, uint8_t* with_host) {
    RB_REQUIRE(d != nullptr, "NULL argument");
    RB_REQUIRE(generation >= 0 && generation < d->gen && generation >= d->gen - d->G + (d->k > 0 ? 1 : 0), "generation is not in the ring any more");
    DeviceGuard guard(d->device);
    const size_t rows = (size_t)EP * d->n, r0 = (size_t)(generation % d->G) * rows;
    RB_CUDA(cudaDeviceSynchronize());
    if (ob_host) RB_CUDA(cudaMemcpy(ob_host, d->ob + r0 * 11, rows * 11 * sizeof(float), cudaMemcpyDeviceToHost));
    if (rew_host) RB_CUDA(cudaMemcpy(rew_host, d->rew + r0, rows * sizeof(float), cudaMemcpyDeviceToHost));
    if (t_host) RB_CUDA(cudaMemcpy(t_host, d->t + r0, rows * sizeof(float4), cudaMemcpyDeviceToHost));
    if (s_host) RB_CUDA(cudaMemcpy(s_host, d->s + r0, rows * sizeof(float4), cudaMemcpyDeviceToHost));
    if (with_host) RB_CUDA(cudaMemcpy(with_host, d->with + r0, rows, cudaMemcpyDeviceToHost));
    return RB_OK;
}

// Whole-ring snapshot for an exact resume of the loops that train from the Dataset (lstm_train.py:86-87,102-107 restore every episode; here the
// replay buffer comes back too, so the restored run draws the same windows): all rows of the ring + the cursor.  Synchronises.
int64_t rb_dataset_ring_rows(const rb_dataset* d) { return d ? d->G * EP * d->n : 0; }
int rb_dataset_save_host(rb_dataset* d, float* ob_host, float* rew_host, float* t_host, float* s_host, uint8_t* with_host, int* step_out,
                         int64_t* generations_out) {
    RB_REQUIRE(d && ob_host && rew_host && t_host && s_host && with_host && step_out && generations_out, "NULL argument");
    DeviceGuard guard(d->device);
    const size_t rows = (size_t)rb_dataset_ring_rows(d);
    RB_CUDA(cudaDeviceSynchronize());
    RB_CUDA(cudaMemcpy(ob_host, d->ob, rows * 11 * sizeof(float), cudaMemcpyDeviceToHost));
    RB_CUDA(cudaMemcpy(rew_host, d->rew, rows * sizeof(float), cudaMemcpyDeviceToHost));
    RB_CUDA(cudaMemcpy(t_host, d->t, rows * sizeof(float4), cudaMemcpyDeviceToHost));
    RB_CUDA(cudaMemcpy(s_host, d->s, rows * sizeof(float4), cudaMemcpyDeviceToHost));
    RB_CUDA(cudaMemcpy(with_host, d->with, rows, cudaMemcpyDeviceToHost));
    *step_out = d->k; *generations_out = d->gen;
    return RB_OK;
}
int rb_dataset_load_host(rb_dataset* d, const float* ob_host, const float* rew_host, const float* t_host, const float* s_host,
                         const uint8_t* with_host, int step, int64_t generations) {
    RB_REQUIRE(d && ob_host && rew_host && t_host && s_host && with_host, "NULL argument");
    RB_REQUIRE(step >= 0 && step <= EP && generations >= 0, "bad cursor");
    DeviceGuard guard(d->device);
    const size_t rows = (size_t)rb_dataset_ring_rows(d);
    RB_CUDA(cudaDeviceSynchronize());
    RB_CUDA(cudaMemcpy(d->ob, ob_host, rows * 11 * sizeof(float), cudaMemcpyHostToDevice));
    RB_CUDA(cudaMemcpy(d->rew, rew_host, rows * sizeof(float), cudaMemcpyHostToDevice));
    RB_CUDA(cudaMemcpy(d->t, t_host, rows * sizeof(float4), cudaMemcpyHostToDevice));
    RB_CUDA(cudaMemcpy(d->s, s_host, rows * sizeof(float4), cudaMemcpyHostToDevice));
    RB_CUDA(cudaMemcpy(d->with, with_host, rows, cudaMemcpyHostToDevice));
    d->k = step; d->gen = generations;
    return RB_OK;
}

int rb_dataset_training_batch(rb_dataset* d, uint64_t seed, uint32_t draw, int B, int T, float* ob_out, float* t_out, float* prev_out, float* prew_out,
                              int32_t* episodes_out, int32_t* start_out, void* stream) {
    RB_REQUIRE(d && ob_out && t_out && prev_out && prew_out, "NULL argument");
    RB_REQUIRE(B > 0 && T > 0 && T <= EP, "bad batch / window size");
    const int64_t avail = rb_dataset_num_available(d) / d->n;       // complete generations not being overwritten
    RB_REQUIRE(avail > 0, "no complete episode in memory");
    RB_REQUIRE(avail * d->n < ((int64_t)1 << 31), "too many episodes for 32-bit sampling");
    k_ds_training_batch<<<(unsigned)((T * B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*d, avail, (uint32_t)seed, (uint32_t)(seed >> 32), draw, B, T,
                                                                                          ob_out, (float4*)t_out, (float4*)prev_out, prew_out,
                                                                                          episodes_out, start_out);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

int rb_dataset_test_batch(rb_dataset* d, const float* ob_cur, int T, float* ob_out, float* prev_out, float* prew_out, void* stream) {
    RB_REQUIRE(d && ob_cur && ob_out, "NULL argument");
    RB_REQUIRE(T > 0 && T <= EP, "bad window size");
    const int64_t total = (int64_t)T * d->n;
    k_ds_test_batch<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*d, d->gen % d->G, d->k, T, ob_cur, ob_out, (float4*)prev_out,
                                                                                      prew_out);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
}

}  // extern "C"
