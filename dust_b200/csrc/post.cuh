// sm_100a device code, part 5: the post chain behind dust_trace:
//   exposure_histogram_kernel   assets/shaders/auto_exposure.comp:22-72
//   exposure_average_kernel     assets/shaders/auto_exposure_avg.comp:19-51
//   tone_map_kernel             assets/shaders/tone_map.comp:36-219
// All three are streaming passes over the frame (8-16 bytes per pixel each way): HBM-bound, grid-stride
// with a multiple of the SM count, 128-bit loads.  The histogram keeps its 256 bins in shared memory
// per CTA and merges with one global atomic per bin.  Bins are integers: bit-exact against the checker
// (log2 is the fixed mul/add/div expression post_log2, identical on the CPU).
#pragma once
#include "trace_math.cuh"

namespace dust {

struct DHistogram { uint32_t histogram[256]; float avg; };                         // auto_exposure.playout

__device__ __forceinline__ float post_log2(float x)
{
    const uint32_t bits = __float_as_uint(x);
    int e = (int)(bits >> 23) - 127;
    float m = __uint_as_float((bits & 0x7FFFFFu) | 0x3F800000u);
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    const float s = (m - 1.0f) / (m + 1.0f), s2 = s * s;
    const float p = 1.0f + s2 * (0.333333333f + s2 * (0.2f + s2 * (0.142857143f + s2 * 0.111111111f)));
    return (float)e + (s * 2.88539008f) * p;
}

// _NRD_YCoCgToLinear of one RGBA16F pixel (headers/nrd.glsl:110-121)
__device__ __forceinline__ f3 unpack_illuminance(uint2 px)
{
    const float2 yc = __half22float2(*reinterpret_cast<const __half2 *>(&px.x));
    const float cg = __low2float(*reinterpret_cast<const __half2 *>(&px.y));
    const float t = yc.x - cg;
    return mk3(fmaxf(t + yc.y, 0.0f), fmaxf(yc.x + cg, 0.0f), fmaxf(t - yc.y, 0.0f));
}

__device__ __forceinline__ uint32_t color_to_bin(f3 c, float min_log_lum, float inv_range)
{
    const float lum = (c.x * 0.299f + c.y * 0.587f) + c.z * 0.114f;
    if (lum < 0.005f) return 0u;
    const float v = (post_log2(lum) - min_log_lum) * inv_range;
    return (uint32_t)(fminf(fmaxf(v, 0.0f), 1.0f) * 254.0f + 1.0f);
}

__global__ void __launch_bounds__(256) exposure_histogram_kernel(const uint2 *__restrict__ img, uint64_t n_pixels, float min_log_lum,
                                                                 float inv_range, DHistogram *__restrict__ h)
{
    __shared__ uint32_t bins[256];
    bins[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t n_pairs = n_pixels >> 1, stride = (uint64_t)gridDim.x * 256u;
    for (uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x; i < n_pairs; i += stride) {
        const uint4 two = __ldg(reinterpret_cast<const uint4 *>(img) + i);          // two pixels per 128-bit load
        atomicAdd(&bins[color_to_bin(unpack_illuminance(make_uint2(two.x, two.y)), min_log_lum, inv_range)], 1u);
        atomicAdd(&bins[color_to_bin(unpack_illuminance(make_uint2(two.z, two.w)), min_log_lum, inv_range)], 1u);
    }
    if ((n_pixels & 1u) && blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&bins[color_to_bin(unpack_illuminance(__ldg(&img[n_pixels - 1])), min_log_lum, inv_range)], 1u);
    __syncthreads();
    if (bins[threadIdx.x]) atomicAdd(&h->histogram[threadIdx.x], bins[threadIdx.x]);
}

// one CTA of 256 threads; resets the histogram like auto_exposure_avg.comp:26
__global__ void __launch_bounds__(256) exposure_average_kernel(DHistogram *h, uint32_t num_pixels, float min_log_lum, float range, float time_coeff)
{
    __shared__ uint32_t sum[256];
    sum[threadIdx.x] = h->histogram[threadIdx.x] * threadIdx.x;
    __syncthreads();
    h->histogram[threadIdx.x] = 0;
    for (uint32_t cutoff = 128; cutoff > 0; cutoff >>= 1) {
        if (threadIdx.x < cutoff) sum[threadIdx.x] += sum[threadIdx.x + cutoff];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float wla = ((float)sum[0] / fmaxf((float)num_pixels, 1.0f)) - 1.0f;
        const float lum = exp2f(((wla / 254.0f) * range) + min_log_lum);
        const float last = h->avg;
        h->avg = last + (lum - last) * time_coeff;
    }
}

struct ToneMapParams { float csc[9]; uint32_t oetf; uint32_t dst_format; };

__device__ __forceinline__ float rrt_odt(float v)
{
    const float a = v * (v + 0.0245786f) - 0.000090537f;
    const float b = v * (0.983729f * v + 0.4329510f) + 0.238081f;
    return a / b;
}
__device__ __forceinline__ float dotc(f3 v, float a, float b, float c) { return (v.x * a + v.y * b) + v.z * c; }
__device__ __forceinline__ float lin2srgb(float c) { return c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f; }
__device__ __forceinline__ float srgb2lin(float c) { return c < 0.04045f ? c / 12.92f : powf(fabsf(c + 0.055f) / 1.055f, 2.4f); }   // color.glsl:1-5

__device__ float oetf1(float c, uint32_t f)                                        // tone_map.comp:70-167
{
    switch (f) {
    case 1: return lin2srgb(c);
    case 2: return c <= -0.0031308f ? -1.055f * powf(-c, 1.0f / 2.4f) + 0.055f : lin2srgb(c);
    case 3: return powf(c / 52.37f, 1.0f / 2.6f);
    case 4: return c < 0.0030186f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
    case 5: return c < 0.0181f ? 4.5f * c : 1.0993f * powf(c, 0.45f) - (1.0993f - 1.0f);
    case 6: {
        const float m1 = 2610.0f / 16384.0f, m2 = (2523.0f / 4096.0f) * 128.0f, c2 = (2413.0f / 4096.0f) * 32.0f, c3 = (2392.0f / 4096.0f) * 32.0f;
        const float c1 = c3 - c2 + 1.0f;
        const float Lm = powf(c, m1);
        return powf((c1 + c2 * Lm) / (1.0f + c3 * Lm), m2);
    }
    case 7: {
        const float a = 0.17883277f, b = 1.0f - 4.0f * a, cc = 0.55991073f;
        return c < (1.0f / 12.0f) ? sqrtf(3.0f * c) : a * logf(12.0f * c - b) + cc;
    }
    case 8: return powf(c, 256.0f / 563.0f);
    default: return c;
    }
}

__global__ void __launch_bounds__(256) tone_map_kernel(const uint2 *__restrict__ img, const uint32_t *__restrict__ albedo, uint64_t n_pixels,
                                                       const DHistogram *__restrict__ h, const __grid_constant__ ToneMapParams p, void *__restrict__ dst)
{
    float exposure = 1.0f / (9.6f * h->avg);                                       // tone_map.comp:196-197
    exposure = exposure * 9.6f;
    const uint64_t stride = (uint64_t)gridDim.x * 256u;
    for (uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x; i < n_pixels; i += stride) {
        const f3 in = unpack_illuminance(__ldg(&img[i]));
        const uint32_t a = __ldg(&albedo[i]);
        const f3 al = mk3((float)((a >> 20) & 1023u) / 1023.0f, (float)((a >> 10) & 1023u) / 1023.0f, (float)(a & 1023u) / 1023.0f);
        f3 s = acescg2srgb(in);
        s = mk3(s.x * srgb2lin(al.x), s.y * srgb2lin(al.y), s.z * srgb2lin(al.z));
        f3 m = srgb2acescg(s);
        m = mk3(m.x * exposure, m.y * exposure, m.z * exposure);
        const f3 v = mat3_mul(p.csc, m);                                           // :202-209
        const f3 r = mk3(rrt_odt(dotc(v, 0.59719f, 0.35458f, 0.04823f)), rrt_odt(dotc(v, 0.07600f, 0.90834f, 0.01566f)),
                         rrt_odt(dotc(v, 0.02840f, 0.13383f, 0.83777f)));           // ACESFitted, :42-68
        const f3 o = mk3(dotc(r, 1.60475f, -0.53108f, -0.07367f), dotc(r, -0.10208f, 1.10813f, -0.00605f), dotc(r, -0.00327f, -0.07276f, 1.07602f));
        const float x = oetf1(o.x, p.oetf), y = oetf1(o.y, p.oetf), z = oetf1(o.z, p.oetf);
        if (p.dst_format == DUST_FMT_RGBA32F) static_cast<float4 *>(dst)[i] = make_float4(x, y, z, 1.0f);
        else {
            const __half2 lo = __floats2half2_rn(x, y), hi = __floats2half2_rn(z, 1.0f);
            static_cast<uint2 *>(dst)[i] = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
        }
    }
}

}  // namespace dust
