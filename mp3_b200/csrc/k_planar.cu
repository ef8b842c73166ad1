// k_planar.cu -- planar copy of the decoded batch (output formats).  The decoder
// writes interleaved PCM (what an audio device wants); tensor-style consumers want [channel][sample].
// One pass over the arena: stream i's samples go from (n, c) at pcm_offset + n * nch + c to
// pcm_offset + c * samples + n.  Pure data movement: HBM-bound, 2 x the arena bytes.
// A CTA takes a tile of 2048 frames of one stereo stream: 16-byte loads of interleaved frames, the
// de-interleave in registers, 16-byte stores to each plane.  Mono streams are a straight copy.
// No reference code exists for this step (/root/reference/README.md:1-84).
#include "kernels.h"
#include "mp3b.h"

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
k_planar(const T *__restrict__ in, T *__restrict__ out, const L3PlanarJob *__restrict__ jobs, const uint32_t *__restrict__ tile_job,
         const uint32_t *__restrict__ tile_first)
{
    constexpr int V = 16 / (int)sizeof(T); // elements per 16-byte vector
    const L3PlanarJob jb = jobs[tile_job[blockIdx.x]];
    const long long n0 = (long long)(blockIdx.x - tile_first[tile_job[blockIdx.x]]) * 2048;
    const long long n1 = n0 + 2048 < jb.samples ? n0 + 2048 : jb.samples;
    const T *x = in + jb.off;
    T *y = out + jb.off;
    if (jb.channels == 1) {
        for (long long n = n0 + threadIdx.x; n < n1; n += 256) y[n] = x[n];
        return;
    }
    T *y0 = y, *y1 = y + jb.samples;
    // vector path when the stream's base and plane starts are 16-byte aligned and the tile is whole
    const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y0) | reinterpret_cast<uintptr_t>(y1)) & 15) == 0 &&
                    (n0 % V) == 0;
    long long n = n0;
    if (al) {
        const long long nv = (n1 - n0) / V; // groups of V frames: two input vectors -> one vector per plane
        for (long long g = threadIdx.x; g < nv; g += 256) {
            const long long f = n0 + g * V;
            const uint4 a = *reinterpret_cast<const uint4 *>(x + f * 2), b = *reinterpret_cast<const uint4 *>(x + f * 2 + V);
            uint4 l, r;
            if (sizeof(T) == 2) { // 8 frames: a = L0 R0 L1 R1 L2 R2 L3 R3, b = L4 R4 ... R7
                l.x = __byte_perm(a.x, a.y, 0x5410); r.x = __byte_perm(a.x, a.y, 0x7632);
                l.y = __byte_perm(a.z, a.w, 0x5410); r.y = __byte_perm(a.z, a.w, 0x7632);
                l.z = __byte_perm(b.x, b.y, 0x5410); r.z = __byte_perm(b.x, b.y, 0x7632);
                l.w = __byte_perm(b.z, b.w, 0x5410); r.w = __byte_perm(b.z, b.w, 0x7632);
            } else {              // 4 frames: a = L0 R0 L1 R1, b = L2 R2 L3 R3
                l = make_uint4(a.x, a.z, b.x, b.z);
                r = make_uint4(a.y, a.w, b.y, b.w);
            }
            *reinterpret_cast<uint4 *>(y0 + f) = l;
            *reinterpret_cast<uint4 *>(y1 + f) = r;
        }
        n = n0 + nv * V;
    }
    for (long long k = n + threadIdx.x; k < n1; k += 256) {
        y0[k] = x[k * 2];
        y1[k] = x[k * 2 + 1];
    }
}

} // namespace

void l3_launch_planar(const void *in, void *out, int pcm_format, const L3PlanarJob *jobs, const uint32_t *tile_job,
                      const uint32_t *tile_first, uint32_t ntiles, cudaStream_t st)
{
    if (!ntiles) return;
    if (pcm_format == MP3B_PCM_S16)
        k_planar<int16_t><<<ntiles, 256, 0, st>>>(static_cast<const int16_t *>(in), static_cast<int16_t *>(out), jobs, tile_job,
                                                  tile_first);
    else
        k_planar<float><<<ntiles, 256, 0, st>>>(static_cast<const float *>(in), static_cast<float *>(out), jobs, tile_job,
                                                tile_first);
}
