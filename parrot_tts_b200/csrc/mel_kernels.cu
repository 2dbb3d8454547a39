// mel_kernels.cu - log-mel spectrogram of generated audio on the device (validation metric beside the 1e-3 sample tolerance).
// Replaces reference utils/vocoder/dataset.py:43-69 (mel_spectrogram): reflect padding by (n_fft - hop) / 2, STFT with a periodic
// Hann window (center = False), magnitude sqrt(re^2 + im^2 + 1e-9), projection on the mel basis, log(clamp(., 1e-5))
// (spectral_normalize_torch, dataset.py:85-95).  One block per (frame, item): the windowed frame goes through an in-place radix-2
// FFT in shared memory (twiddles from sincospif: exact arguments, ~1 ulp), the 1 + n_fft / 2 magnitudes stay in shared memory and
// every warp reduces whole rows of the mel basis against them (coalesced reads of the basis, warp-shuffle sum).
#include "pb_common.cuh"

namespace pb {

template <int NFFT, int LOG2>
__global__ void __launch_bounds__(256)
mel_spectrogram_kernel(const float* __restrict__ y, int64_t N, int frames, int hop, int pad, const float* __restrict__ mel_basis,
                       int n_mels, float* __restrict__ out) {
    __shared__ float2 buf[NFFT];
    __shared__ float mag[NFFT / 2 + 1];
    const int f = blockIdx.x, b = blockIdx.y;
    const float* yb = y + (size_t)b * N;
    for (int i = threadIdx.x; i < NFFT; i += blockDim.x) {
        int64_t idx = (int64_t)f * hop + i - pad;
        if (idx < 0) idx = -idx;                          // reflect (no edge repeat), as F.pad(mode='reflect')
        if (idx >= N) idx = 2 * (N - 1) - idx;
        const float w = 0.5f - 0.5f * cospif(2.0f * (float)i / (float)NFFT);      // torch.hann_window(win) (periodic)
        buf[__brev((unsigned)i) >> (32 - LOG2)] = make_float2(yb[idx] * w, 0.f);
    }
    __syncthreads();
#pragma unroll 1
    for (int s = 0; s < LOG2; ++s) {
        const int half = 1 << s;
        for (int j = threadIdx.x; j < NFFT / 2; j += blockDim.x) {
            const int pos = j & (half - 1);
            const int i0 = ((j >> s) << (s + 1)) + pos, i1 = i0 + half;
            float sn, cs;
            sincospif(-(float)pos / (float)half, &sn, &cs);          // exp(-2 pi i pos / (2 half))
            const float2 a = buf[i0], c = buf[i1];
            const float2 t = make_float2(c.x * cs - c.y * sn, c.x * sn + c.y * cs);
            buf[i0] = make_float2(a.x + t.x, a.y + t.y);
            buf[i1] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k <= NFFT / 2; k += blockDim.x) {
        const float2 v = buf[k];
        mag[k] = sqrtf(v.x * v.x + v.y * v.y + 1e-9f);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NB = NFFT / 2 + 1;
    for (int m = warp; m < n_mels; m += (int)(blockDim.x >> 5)) {
        const float* row = mel_basis + (size_t)m * NB;
        float acc = 0.f;
        for (int k = lane; k < NB; k += 32) acc = fmaf(row[k], mag[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) out[((size_t)b * n_mels + m) * frames + f] = logf(fmaxf(acc, 1e-5f));
    }
}

int launch_mel_spectrogram(const float* y, int B, int64_t N, const float* mel_basis, int n_fft, int hop, int win, int n_mels, float* out,
                           int* frames_out, cudaStream_t st) {
    PB_REQUIRE(win == n_fft, PB_ERR_INVALID, "mel_spectrogram: win_size %d != n_fft %d is not supported (the reference uses 1024 / 1024)", win, n_fft);
    PB_REQUIRE(hop > 0 && hop <= n_fft && (n_fft - hop) % 2 == 0, PB_ERR_INVALID, "mel_spectrogram: hop %d", hop);
    const int pad = (n_fft - hop) / 2;
    PB_REQUIRE(B > 0 && B <= 65535 && N > pad, PB_ERR_INVALID, "mel_spectrogram: need more than %d samples per item for the reflect padding (got %lld)", pad, (long long)N);
    const int64_t frames64 = (N + 2 * pad - n_fft) / hop + 1;
    PB_REQUIRE(frames64 > 0 && frames64 < (1LL << 31), PB_ERR_INVALID, "mel_spectrogram: bad frame count");
    const int frames = (int)frames64;
    if (frames_out) *frames_out = frames;
    if (!out) return PB_OK;                      // size query
    dim3 grid((unsigned)frames, (unsigned)B);
    switch (n_fft) {
        case 256: mel_spectrogram_kernel<256, 8><<<grid, 256, 0, st>>>(y, N, frames, hop, pad, mel_basis, n_mels, out); break;
        case 512: mel_spectrogram_kernel<512, 9><<<grid, 256, 0, st>>>(y, N, frames, hop, pad, mel_basis, n_mels, out); break;
        case 1024: mel_spectrogram_kernel<1024, 10><<<grid, 256, 0, st>>>(y, N, frames, hop, pad, mel_basis, n_mels, out); break;
        case 2048: mel_spectrogram_kernel<2048, 11><<<grid, 256, 0, st>>>(y, N, frames, hop, pad, mel_basis, n_mels, out); break;
        default: set_error("mel_spectrogram: n_fft=%d unsupported (256, 512, 1024 or 2048)", n_fft); return PB_ERR_INVALID;
    }
    PB_COUNT_LAUNCH();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace pb

// mel_spectrogram (reference utils/vocoder/dataset.py:43-69) for y f32 [B, N] on the device -> out f32 [B, n_mels, frames],
// frames = (N + (n_fft - hop) - n_fft) / hop + 1.  out == NULL: only *frames is written (size query, no launch).
extern "C" int pb_mel_spectrogram(const float* y, int32_t B, int64_t N, const float* mel_basis, int32_t n_fft, int32_t hop, int32_t win,
                                  int32_t n_mels, float* out, int32_t* frames, void* stream) {
    PB_REQUIRE(frames != nullptr && (out == nullptr || (y != nullptr && mel_basis != nullptr)) && n_mels > 0, PB_ERR_INVALID,
               "pb_mel_spectrogram: null argument");
    int fr = 0;
    const int rc = pb::launch_mel_spectrogram(y, B, N, mel_basis, n_fft, hop, win, n_mels, out, &fr, (cudaStream_t)stream);
    *frames = fr;
    return rc;
}
