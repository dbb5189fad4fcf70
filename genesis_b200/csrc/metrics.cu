// genesis_b200 -- segmentation metrics of the evaluation loop on the device: adjusted Rand index (all pixels / foreground
// only) and segmentation covering (unweighted / size-weighted, with / without background) per image, from ONE confusion
// matrix per image.  Replaces the per-image numpy / sklearn loops of the reference's utils/misc.py:101-114 (average_ari,
// sklearn.metrics.adjusted_rand_score) and :173-235 (average_segcover, iou_binary), called from train.py:536-546, and the
// argmax over slots of train.py:541 / genesisv2_config.py:187-188.
#include "common.cuh"

namespace {

constexpr int SM_MAXA = 32;     // ground-truth instance labels 0 .. 31 (0 = background)
constexpr int SM_MAXK = 16;     // predicted slots

// grid = B images, 256 threads.  log_m [K,B,P] (slot-major, as stats.log_m_k stacked) or pred [B,P] int64; inst [B,P] int64.
// out [B][8] doubles: ari, ari_fg, msc, msc_fg, msc_scaled, msc_fg_scaled, #gt labels present, #pixels counted.
__global__ void __launch_bounds__(256) seg_metrics_kernel(const float* __restrict__ log_m, const long long* __restrict__ pred,
                                                          const long long* __restrict__ inst, long long* __restrict__ seg_out,
                                                          double* __restrict__ out, int B, int P, int K) {
    __shared__ unsigned int conf[SM_MAXA][SM_MAXK];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < SM_MAXA * SM_MAXK; i += blockDim.x) (&conf[0][0])[i] = 0u;
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        int j;
        if (log_m) {            // np.argmax over slots: first maximum wins
            float best = __ldg(log_m + (long)b * P + p);
            j = 0;
            for (int k = 1; k < K; ++k) {
                const float v = __ldg(log_m + ((long)k * B + b) * P + p);
                if (v > best) { best = v; j = k; }
            }
        } else {
            j = (int)pred[(long)b * P + p];
        }
        if (seg_out) seg_out[(long)b * P + p] = j;
        const long long a = inst[(long)b * P + p];
        if (a >= 0 && a < SM_MAXA && j >= 0 && j < SM_MAXK) atomicAdd(&conf[(int)a][j], 1u);
    }
    __syncthreads();
    if (threadIdx.x >= 2) return;
    // thread 0: all labels; thread 1: foreground only (ground-truth label > 0)
    const int a0 = threadIdx.x;
    double n = 0.0, sumsq = 0.0, dot_k = 0.0, dot_c = 0.0;
    double colsum[SM_MAXK], colall[SM_MAXK];
    for (int j = 0; j < SM_MAXK; ++j) { colsum[j] = 0.0; colall[j] = 0.0; }
    for (int a = 0; a < SM_MAXA; ++a)
        for (int j = 0; j < SM_MAXK; ++j) {
            colall[j] += conf[a][j];                       // covering: predicted masks are NOT restricted to the foreground
            if (a >= a0) colsum[j] += conf[a][j];
        }
    // adjusted Rand index through sklearn's pair confusion matrix (sklearn/metrics/cluster/_supervised.py)
    for (int a = a0; a < SM_MAXA; ++a) {
        double row = 0.0;
        for (int j = 0; j < SM_MAXK; ++j) row += conf[a][j];
        for (int j = 0; j < SM_MAXK; ++j) {
            const double c = conf[a][j];
            sumsq += c * c; dot_k += c * colsum[j]; dot_c += c * row;
        }
        n += row;
    }
    const double tp = sumsq - n, fp = dot_k - sumsq, fn = dot_c - sumsq, tn = n * n - fp - fn - sumsq;
    double ari = 1.0;
    if (!(fn == 0.0 && fp == 0.0)) ari = 2.0 * (tp * tn - fn * fp) / ((tp + fn) * (fn + tn) + (tp + fp) * (fp + tn));
    // segmentation covering of the ground truth by the prediction (utils/misc.py:173-235)
    double mean_scores = 0.0, scaled = 0.0, scaling = 0.0;
    int nlab = 0;
    for (int a = a0; a < SM_MAXA; ++a) {
        double row = 0.0;
        for (int j = 0; j < SM_MAXK; ++j) row += conf[a][j];
        if (row == 0.0) continue;
        double best = 0.0;
        for (int j = 0; j < SM_MAXK; ++j) {
            const double inter = conf[a][j], uni = row + colall[j] - inter;
            if (uni > 0.0 && inter / uni > best) best = inter / uni;
        }
        mean_scores += best; scaled += row * best; scaling += row; ++nlab;
    }
    const double msc = mean_scores / (nlab > 0 ? nlab : 1), mscs = scaled / (scaling > 0.0 ? scaling : 1.0);
    double* o = out + (long)b * 8;
    if (a0 == 0) { o[0] = ari; o[2] = msc; o[4] = mscs; o[6] = nlab; o[7] = n; }
    else { o[1] = ari; o[3] = msc; o[5] = mscs; }
}

}  // namespace

extern "C" int g2_seg_metrics(const float* log_m, const int64_t* pred, const int64_t* inst, int64_t* seg_out, double* out, int B,
                              int P, int K, cudaStream_t stream) {
    G2_CHECK_ARG((log_m != nullptr) != (pred != nullptr));
    G2_CHECK_ARG(inst && out && B > 0 && P > 0 && K >= 1 && K <= SM_MAXK);
    seg_metrics_kernel<<<B, 256, 0, stream>>>(log_m, reinterpret_cast<const long long*>(pred), reinterpret_cast<const long long*>(inst),
                                              reinterpret_cast<long long*>(seg_out), out, B, P, K);
    G2_LAUNCH_RET();
}
