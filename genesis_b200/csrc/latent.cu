// genesis_b200 -- fused elementwise kernels for the latent path: the LSTM cell (posterior modules/attention.py:94-97 and
// prior models/genesis_config.py:301-305, one step of torch.nn.LSTM after the two gate GEMMs), the Gaussian head
// (to_sigma + rsample: modules/blocks.py:22-23, component_vae.py:67-69, attention.py:98-103) and the Monte-Carlo KL
// (models/genesis_config.py:328-336, utils/misc.py:254-255).  Each replaces 8-25 ATen launches on [B,64..512] tensors
// with one; the step's critical path runs ~230 of those launches back to back (profiles/r01_launches_final_summary.txt).
// All kernels: one thread per element (or per row for the KL sums), fp32, no workspace.
#include "common.cuh"

namespace {

__device__ __forceinline__ float softplusf(float x) { return x > 20.f ? x : log1pf(expf(x)); }     // torch threshold = 20

// gates = gx + gh, laid out [B, 4H] in torch's order i, f, g, o
__global__ void lstm_cell_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ gh, const float* __restrict__ c_prev,
                                     float* __restrict__ h, float* __restrict__ c, int B, int H) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    const long base = (long)b * 4 * H + j;
    const float gi = gx[base] + gh[base], gf = gx[base + H] + gh[base + H];
    const float gg = gx[base + 2 * H] + gh[base + 2 * H], go = gx[base + 3 * H] + gh[base + 3 * H];
    const float i = g2_sigmoidf(gi), f = g2_sigmoidf(gf), g = tanhf(gg), o = g2_sigmoidf(go);
    const float cn = (c_prev ? f * c_prev[idx] : 0.f) + i * g;
    c[idx] = cn;
    h[idx] = o * tanhf(cn);
}

// dgates [B,4H] (gradient of both gx and gh), dc_prev [B,H] from dh, dc (either may be NULL = zero)
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ gx, const float* __restrict__ gh, const float* __restrict__ c_prev,
                                     const float* __restrict__ c, const float* __restrict__ dh, const float* __restrict__ dc,
                                     float* __restrict__ dgates, float* __restrict__ dc_prev, int B, int H) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, j = idx - b * H;
    const long base = (long)b * 4 * H + j;
    const float i = g2_sigmoidf(gx[base] + gh[base]), f = g2_sigmoidf(gx[base + H] + gh[base + H]);
    const float g = tanhf(gx[base + 2 * H] + gh[base + 2 * H]), o = g2_sigmoidf(gx[base + 3 * H] + gh[base + 3 * H]);
    const float tc = tanhf(c[idx]);
    const float dhv = dh ? dh[idx] : 0.f;
    const float dcv = (dc ? dc[idx] : 0.f) + dhv * o * (1.f - tc * tc);
    const float cp = c_prev ? c_prev[idx] : 0.f;
    dgates[base] = dcv * g * i * (1.f - i);
    dgates[base + H] = dcv * cp * f * (1.f - f);
    dgates[base + 2 * H] = dcv * i * (1.f - g * g);
    dgates[base + 3 * H] = dhv * tc * o * (1.f - o);
    if (dc_prev) dc_prev[idx] = dcv * f;
}

// lo [B, 2D] = (mu | raw), the un-chunked output of the head's Linear.  sigma = softplus(raw + 0.5) + 1e-8 (blocks.to_sigma);
// z = mu + sigma * eps;  mu is also written out contiguously for the KL kernel.
__global__ void gauss_head_fwd_kernel(const float* __restrict__ lo, const float* __restrict__ eps, float* __restrict__ z,
                                      float* __restrict__ mu, float* __restrict__ sigma, int B, int D) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * D) return;
    const long b = i / D, d = i - b * D;
    const float m = lo[b * 2 * D + d];
    const float s = softplusf(lo[b * 2 * D + D + d] + 0.5f) + 1e-8f;
    mu[i] = m;
    sigma[i] = s;
    z[i] = m + s * eps[i];
}

// dlo[:, :D] = dz + dmu;  dlo[:, D:] = (dz * eps + dsigma) * sigmoid(raw + 0.5)     (each of dz / dmu / dsigma may be NULL)
__global__ void gauss_head_bwd_kernel(const float* __restrict__ lo, const float* __restrict__ eps, const float* __restrict__ dz,
                                      const float* __restrict__ dmu, const float* __restrict__ dsigma, float* __restrict__ dlo,
                                      int B, int D) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * D) return;
    const long b = i / D, d = i - b * D;
    const float a = dz ? dz[i] : 0.f, s = dsigma ? dsigma[i] : 0.f;
    const float x = lo[b * 2 * D + D + d] + 0.5f;
    dlo[b * 2 * D + d] = a + (dmu ? dmu[i] : 0.f);
    dlo[b * 2 * D + D + d] = (a * eps[i] + s) * (x > 20.f ? 1.f : g2_sigmoidf(x));
}

// Prior head: lo [B, 2D] = (a | b) -> pmu = tanh(a) (or a when use_tanh = 0: Genesis.sample, genesis_config.py:359),
// psigma = sigmoid(b + 4) + 1e-4 (blocks.to_prior_sigma, modules/blocks.py:28-34)
__global__ void prior_head_fwd_kernel(const float* __restrict__ lo, float* __restrict__ pmu, float* __restrict__ psigma, int B, int D,
                                      int use_tanh) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * D) return;
    const long b = i / D, d = i - b * D;
    const float a = lo[b * 2 * D + d];
    pmu[i] = use_tanh ? tanhf(a) : a;
    psigma[i] = g2_sigmoidf(lo[b * 2 * D + D + d] + 4.f) + 1e-4f;
}

__global__ void prior_head_bwd_kernel(const float* __restrict__ pmu, const float* __restrict__ psigma, const float* __restrict__ dpmu,
                                      const float* __restrict__ dpsigma, float* __restrict__ dlo, int B, int D, int use_tanh) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * D) return;
    const long b = i / D, d = i - b * D;
    const float m = pmu[i], sg = psigma[i] - 1e-4f;
    dlo[b * 2 * D + d] = (dpmu ? dpmu[i] : 0.f) * (use_tanh ? 1.f - m * m : 1.f);
    dlo[b * 2 * D + D + d] = (dpsigma ? dpsigma[i] : 0.f) * sg * (1.f - sg);
}

// kl[b] = sum_d [ log N(z; mu, sigma) - log N(z; pmu, psigma) ];  pmu == NULL: standard-normal prior.  One warp per row.
__global__ void mc_kl_fwd_kernel(const float* __restrict__ z, const float* __restrict__ mu, const float* __restrict__ sigma,
                                 const float* __restrict__ pmu, const float* __restrict__ psigma, float* __restrict__ kl, int B, int D) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= B) return;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) {
        const long i = (long)row * D + d;
        const float zz = z[i], s = sigma[i], t = (zz - mu[i]) / s;
        float lq = -0.5f * t * t - logf(s), lp;
        if (pmu) { const float ps = psigma[i], u = (zz - pmu[i]) / ps; lp = -0.5f * u * u - logf(ps); }
        else lp = -0.5f * zz * zz;
        acc += lq - lp;                                           // the -0.5 log(2 pi) terms cancel
    }
    acc = g2_warp_sum(acc);
    if (lane == 0) kl[row] = acc;
}

__global__ void mc_kl_bwd_kernel(const float* __restrict__ z, const float* __restrict__ mu, const float* __restrict__ sigma,
                                 const float* __restrict__ pmu, const float* __restrict__ psigma, const float* __restrict__ dkl,
                                 float* __restrict__ dz, float* __restrict__ dmu, float* __restrict__ dsigma,
                                 float* __restrict__ dpmu, float* __restrict__ dpsigma, int B, int D) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * D) return;
    const float g = dkl[i / D];
    const float zz = z[i], s = sigma[i], r = zz - mu[i];
    const float inv = 1.f / (s * s);
    float gz = -r * inv;                                          // d log q / dz
    dmu[i] = g * r * inv;
    dsigma[i] = g * (r * r * inv / s - 1.f / s);
    if (pmu) {
        const float ps = psigma[i], pr = zz - pmu[i], pinv = 1.f / (ps * ps);
        gz += pr * pinv;                                          // - d log p / dz
        dpmu[i] = -g * pr * pinv;
        dpsigma[i] = -g * (pr * pr * pinv / ps - 1.f / ps);
    } else gz += zz;
    dz[i] = g * gz;
}

inline int blocks_for(long n) { return (int)((n + 255) / 256); }

}  // namespace

extern "C" {

int g2_lstm_cell_fwd_f32(const float* gx, const float* gh, const float* c_prev, float* h, float* c, int B, int H, cudaStream_t stream) {
    G2_CHECK_ARG(gx && gh && h && c && B > 0 && H > 0);
    lstm_cell_fwd_kernel<<<blocks_for((long)B * H), 256, 0, stream>>>(gx, gh, c_prev, h, c, B, H);
    G2_LAUNCH_RET();
}

int g2_lstm_cell_bwd_f32(const float* gx, const float* gh, const float* c_prev, const float* c, const float* dh, const float* dc,
                         float* dgates, float* dc_prev, int B, int H, cudaStream_t stream) {
    G2_CHECK_ARG(gx && gh && c && dgates && B > 0 && H > 0);
    lstm_cell_bwd_kernel<<<blocks_for((long)B * H), 256, 0, stream>>>(gx, gh, c_prev, c, dh, dc, dgates, dc_prev, B, H);
    G2_LAUNCH_RET();
}

int g2_gauss_head_fwd_f32(const float* lo, const float* eps, float* z, float* mu, float* sigma, int B, int D, cudaStream_t stream) {
    G2_CHECK_ARG(lo && eps && z && mu && sigma && B > 0 && D > 0);
    gauss_head_fwd_kernel<<<blocks_for((long)B * D), 256, 0, stream>>>(lo, eps, z, mu, sigma, B, D);
    G2_LAUNCH_RET();
}

int g2_gauss_head_bwd_f32(const float* lo, const float* eps, const float* dz, const float* dmu, const float* dsigma, float* dlo,
                          int B, int D, cudaStream_t stream) {
    G2_CHECK_ARG(lo && eps && dlo && B > 0 && D > 0);
    gauss_head_bwd_kernel<<<blocks_for((long)B * D), 256, 0, stream>>>(lo, eps, dz, dmu, dsigma, dlo, B, D);
    G2_LAUNCH_RET();
}

int g2_prior_head_fwd_f32(const float* lo, float* pmu, float* psigma, int B, int D, int use_tanh, cudaStream_t stream) {
    G2_CHECK_ARG(lo && pmu && psigma && B > 0 && D > 0);
    prior_head_fwd_kernel<<<blocks_for((long)B * D), 256, 0, stream>>>(lo, pmu, psigma, B, D, use_tanh);
    G2_LAUNCH_RET();
}

int g2_prior_head_bwd_f32(const float* pmu, const float* psigma, const float* dpmu, const float* dpsigma, float* dlo, int B, int D,
                          int use_tanh, cudaStream_t stream) {
    G2_CHECK_ARG(pmu && psigma && dlo && B > 0 && D > 0);
    prior_head_bwd_kernel<<<blocks_for((long)B * D), 256, 0, stream>>>(pmu, psigma, dpmu, dpsigma, dlo, B, D, use_tanh);
    G2_LAUNCH_RET();
}

int g2_mc_kl_fwd_f32(const float* z, const float* mu, const float* sigma, const float* pmu, const float* psigma, float* kl, int B,
                     int D, cudaStream_t stream) {
    G2_CHECK_ARG(z && mu && sigma && kl && B > 0 && D > 0 && ((pmu != nullptr) == (psigma != nullptr)));
    mc_kl_fwd_kernel<<<blocks_for((long)B * 32), 256, 0, stream>>>(z, mu, sigma, pmu, psigma, kl, B, D);
    G2_LAUNCH_RET();
}

int g2_mc_kl_bwd_f32(const float* z, const float* mu, const float* sigma, const float* pmu, const float* psigma, const float* dkl,
                     float* dz, float* dmu, float* dsigma, float* dpmu, float* dpsigma, int B, int D, cudaStream_t stream) {
    G2_CHECK_ARG(z && mu && sigma && dkl && dz && dmu && dsigma && B > 0 && D > 0);
    G2_CHECK_ARG((pmu != nullptr) == (psigma != nullptr) && (pmu == nullptr || (dpmu && dpsigma)));
    mc_kl_bwd_kernel<<<blocks_for((long)B * D), 256, 0, stream>>>(z, mu, sigma, pmu, psigma, dkl, dz, dmu, dsigma, dpmu, dpsigma, B, D);
    G2_LAUNCH_RET();
}

}  // extern "C"
