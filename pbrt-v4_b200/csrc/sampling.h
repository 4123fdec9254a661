// sampling.h -- hashing, PCG32, ZSobol / Independent samplers and the sampling routines on the path.
//
// Restates (CPU branches, operation for operation):
//   MurmurHash64A / MixBits / Hash(...)        util/hash.h:20-116
//   RNG (PCG32)                                util/rng.h
//   ZSobolSampler, IndependentSampler          samplers.h:225-364, 442-476
//   SobolSample, FastOwenScrambler             util/lowdiscrepancy.h:166-232
//   SampleUniformDiskConcentric, SampleCosineHemisphere, SampleUniformTriangle, SampleLinear,
//   SampleBilinear, BilinearPDF                util/sampling.h:120-160,173-183,325-417
//   SampleSphericalTriangle, InvertSphericalTriangleSample   util/sampling.cpp
//   SampleVisibleWavelengths / PDF             util/sampling.h:163-171
#pragma once
#include <vector>
#include "pmath.h"

namespace bpt {

// ------------------------------------------------------------------ hashing
BPT_HD uint64_t MurmurHash64A(const uint64_t *words, int len, uint64_t seed) {
    // `words` holds the key zero-padded to a multiple of 8 bytes; `len` is the key length in bytes.
    const uint64_t m = 0xc6a4a7935bd1e995ull;
    const int r = 47;
    uint64_t h = seed ^ ((uint64_t)len * m);
    int nfull = len / 8;
    for (int i = 0; i < nfull; ++i) {
        uint64_t k = words[i];
        k *= m; k ^= k >> r; k *= m;
        h ^= k; h *= m;
    }
    int rem = len & 7;
    if (rem) {
        // tail bytes key[0..rem-1] occupy the low bytes of the (little-endian) padded word
        uint64_t t = words[nfull] & ((1ull << (8 * rem)) - 1);
        h ^= t;
        h *= m;
    }
    h ^= h >> r; h *= m; h ^= h >> r;
    return h;
}
BPT_HD uint64_t MixBits(uint64_t v) {
    v ^= (v >> 31); v *= 0x7fb5d329728ea185ull;
    v ^= (v >> 27); v *= 0x81dadef4bc2dd44dull;
    v ^= (v >> 33);
    return v;
}
BPT_HD uint64_t Hash_i32x2(int a, int b) {   // Hash(int, int)
    uint64_t w = (uint64_t)(uint32_t)a | ((uint64_t)(uint32_t)b << 32);
    return MurmurHash64A(&w, 8, 0);
}
BPT_HD uint64_t Hash_f32(float a) { uint64_t w = f2b(a); return MurmurHash64A(&w, 4, 0); }
BPT_HD uint64_t Hash_f32x2(float a, float b) {
    uint64_t w = (uint64_t)f2b(a) | ((uint64_t)f2b(b) << 32);
    return MurmurHash64A(&w, 8, 0);
}
BPT_HD uint64_t Hash_f32x3(float a, float b, float c) {   // Hash(Point3f) / Hash(Vector3f)
    uint64_t w[2] = {(uint64_t)f2b(a) | ((uint64_t)f2b(b) << 32), (uint64_t)f2b(c)};
    return MurmurHash64A(w, 12, 0);
}

// ------------------------------------------------------------------ PCG32 (util/rng.h)
struct RNG {
    uint64_t state, inc;
    BPT_HD RNG() : state(0x853c49e6748fea9bULL), inc(0xda3e39cb94b95bdbULL) {}
    BPT_HD RNG(uint64_t seqIndex, uint64_t offset) { SetSequence(seqIndex, offset); }
    BPT_HD explicit RNG(uint64_t seqIndex) { SetSequence(seqIndex, MixBits(seqIndex)); }
    BPT_HD void SetSequence(uint64_t sequenceIndex, uint64_t seed) {
        state = 0u;
        inc = (sequenceIndex << 1u) | 1u;
        UniformU32();
        state += seed;
        UniformU32();
    }
    BPT_HD uint32_t UniformU32() {
        uint64_t oldstate = state;
        state = oldstate * 0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t)(((oldstate >> 18u) ^ oldstate) >> 27u);
        uint32_t rot = (uint32_t)(oldstate >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    BPT_HD float UniformFloat() { return stdmin(kOneMinusEpsilon, (float)UniformU32() * 0x1p-32f); }
    BPT_HD void Advance(int64_t idelta) {
        uint64_t curMult = 0x5851f42d4c957f2dULL, curPlus = inc, accMult = 1u;
        uint64_t accPlus = 0u, delta = (uint64_t)idelta;
        while (delta > 0) {
            if (delta & 1) { accMult *= curMult; accPlus = accPlus * curMult + curPlus; }
            curPlus = (curMult + 1) * curPlus;
            curMult *= curMult;
            delta /= 2;
        }
        state = accMult * state + accPlus;
    }
};

// ------------------------------------------------------------------ Sobol' + ZSobol
// First two 32-bit Sobol' generator matrices, 52 columns each (util/sobolmatrices.cpp; table data is
// loaded from the data blob by the host and copied into constant memory / this host-side array).
#if defined(__CUDACC__) && defined(BPT_SOBOL_DEVICE_TABLE)
__constant__ uint32_t c_sobol32[2 * 52];   // defined by the one translation unit that samples on the device
__device__ uint32_t g_sobolByteLUT[2 * 4 * 256];   // [dimension][byte position][byte value] partial products
#endif
extern uint32_t h_sobol32[2 * 52];
BPT_HD const uint32_t *SobolTable() {
#if defined(__CUDA_ARCH__) && defined(BPT_SOBOL_DEVICE_TABLE)
    return c_sobol32;
#else
    return h_sobol32;
#endif
}

BPT_HD uint32_t ReverseBits32(uint32_t n) {
#if defined(__CUDA_ARCH__)
    return __brev(n);
#else
    n = (n << 16) | (n >> 16);
    n = ((n & 0x00ff00ff) << 8) | ((n & 0xff00ff00) >> 8);
    n = ((n & 0x0f0f0f0f) << 4) | ((n & 0xf0f0f0f0) >> 4);
    n = ((n & 0x33333333) << 2) | ((n & 0xcccccccc) >> 2);
    n = ((n & 0x55555555) << 1) | ((n & 0xaaaaaaaa) >> 1);
    return n;
#endif
}
BPT_HD uint32_t FastOwenScramble(uint32_t v, uint32_t seed) {
    v = ReverseBits32(v);
    v ^= v * 0x3d20adea;
    v += seed;
    v *= (seed >> 16) | 1;
    v ^= v * 0x05526c56;
    v ^= v * 0x53a22864;
    return ReverseBits32(v);
}
enum RandomizeStrategy { RANDOMIZE_NONE = 0, RANDOMIZE_PERMUTE_DIGITS = 1, RANDOMIZE_FAST_OWEN = 2, RANDOMIZE_OWEN = 3 };

BPT_HD uint32_t OwenScramble(uint32_t v, uint32_t seed) {   // lowdiscrepancy.h OwenScrambler
    if (seed & 1) v ^= 1u << 31;
    for (int b = 1; b < 32; ++b) {
        uint32_t mask = (~0u) << (32 - b);
        if ((uint32_t)MixBits((v & mask) ^ seed) & (1u << b)) v ^= 1u << (31 - b);
    }
    return v;
}

BPT_HD float SobolSampleRand(uint64_t a, int dimension, int randomize, uint32_t hash) {
    const uint32_t *C = SobolTable() + dimension * 52;
    uint32_t v = 0;
#if defined(__CUDA_ARCH__) && defined(BPT_SOBOL_DEVICE_TABLE)
    // the generator-matrix product is linear over GF(2): XOR of four per-byte partial products from a 4 KB table
    // (same value as the bit loop below; the loop only runs for indices >= 2^32)
    const uint32_t *T = g_sobolByteLUT + dimension * 1024;
    uint32_t lo = (uint32_t)a;
    v = __ldg(T + (lo & 255u)) ^ __ldg(T + 256 + ((lo >> 8) & 255u)) ^ __ldg(T + 512 + ((lo >> 16) & 255u)) ^ __ldg(T + 768 + (lo >> 24));
    a >>= 32;
    for (int i = 32; a != 0; a >>= 1, i++)
        if (a & 1) v ^= C[i];
#else
    for (int i = 0; a != 0; a >>= 1, i++)
        if (a & 1) v ^= C[i];
#endif
    if (randomize == RANDOMIZE_PERMUTE_DIGITS) v = hash ^ v;
    else if (randomize == RANDOMIZE_FAST_OWEN) v = FastOwenScramble(v, hash);
    else if (randomize == RANDOMIZE_OWEN) v = OwenScramble(v, hash);
    return stdmin((float)v * 0x1p-32f, kOneMinusEpsilon);
}

BPT_HD uint64_t LeftShift2(uint64_t x) {
    x &= 0xffffffff;
    x = (x ^ (x << 16)) & 0x0000ffff0000ffffull;
    x = (x ^ (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x ^ (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x ^ (x << 2)) & 0x3333333333333333ull;
    x = (x ^ (x << 1)) & 0x5555555555555555ull;
    return x;
}
BPT_HD uint64_t EncodeMorton2(uint32_t x, uint32_t y) { return (LeftShift2(y) << 1) | LeftShift2(x); }

// PermutationElement (util/math.h:726-755): the i-th element of a hashed permutation of [0, l)
BPT_HD int PermutationElement(uint32_t i, uint32_t l, uint32_t p) {
    uint32_t w = l - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p; i *= 0xe170893d; i ^= p >> 16; i ^= (i & w) >> 4; i ^= p >> 8; i *= 0x0929eb3f; i ^= p >> 23; i ^= (i & w) >> 1;
        i *= 1 | p >> 27; i *= 0x6935fa69; i ^= (i & w) >> 11; i *= 0x74dcb303; i ^= (i & w) >> 2; i *= 0x9e501cc3; i ^= (i & w) >> 2;
        i *= 0xc860a3df; i &= w; i ^= i >> 5;
    } while (i >= l);
    return (int)((i + p) % l);
}
// Hash(Point2i pixel, int dimension, int seed) (util/hash.h:99-107: the arguments packed back to back, 16 bytes)
BPT_HD uint64_t Hash_pixel_dim_seed(int px, int py, int dimension, int seed) {
    uint64_t w[2] = {(uint64_t)(uint32_t)px | ((uint64_t)(uint32_t)py << 32), (uint64_t)(uint32_t)dimension | ((uint64_t)(uint32_t)seed << 32)};
    return MurmurHash64A(w, 16, 0);
}

struct SamplerParams {
    int type;                  // 0 = zsobol, 1 = independent, 2 = stratified, 3 = paddedsobol, 4 = sobol, 5 = halton, 6 = pmj02bn
    int xSamples, ySamples, jitter;   // stratified (samplers.h:568-625)
    int randomize;             // RandomizeStrategy
    int seed;
    int log2SamplesPerPixel;
    int nBase4Digits;
    int samplesPerPixel;
    // Optional device table (ZSobol): for every (dimension < prefixDims, pixel of the film) the permuted base-4 digits of the
    // sample index that depend on the pixel only (digits >= nLowDigits).  The hashes that choose their permutations are the same
    // for every sample of the pixel, so they are computed once per frame (k_zsobol_prefix) instead of once per Get1D/Get2D call.
    const uint32_t *prefix;
    int prefixDims, prefixW, prefixH, nLowDigits;
    // SobolSampler (samplers.h:479-565): all 1024 generator matrices and the van der Corput / Sobol' index matrices of
    // SobolIntervalToIndex (util/sobolmatrices.cpp), in global memory; scale = RoundUpPow2(max(full resolution))
    const uint32_t *sobolFull;   // [1024][52]
    const uint64_t *vdc, *vdcInv;   // [25][52]
    int log2Scale, scale;
    // HaltonSampler (samplers.h:33-141, samplers.cpp:32-52): the first 1000 primes, the digit permutations of
    // ComputeRadicalInversePermutations(seed) (one block of nDigits x base 16-bit entries per prime, at permOffset[prime index]),
    // and the constants that take a pixel to its first Halton index
    const int *primes;
    const uint16_t *haltonPerms;
    const uint32_t *haltonPermOffset;
    int baseScales[2], baseExponents[2], multInverse[2];
    // PMJ02BNSampler (samplers.h:367-439, samplers.cpp:173-207): the five 65536-point pmj02bn sets (fixed point), the 48 blue-noise
    // textures [texture][x][y], and the pixel samples of the first set sorted into pixelTileSize^2 pixels
    const uint32_t *pmjSamples;     // [5][65536][2]
    const uint16_t *blueNoise;      // [48][128][128]
    const float *pmjPixelSamples;   // [pixelTileSize^2 * samplesPerPixel][2]
    int pixelTileSize;
};

// BPT_OUTLINE_SAMPLER (device, the translation unit that owns the Sobol tables): ZSobol Get1D / Get2D as ONE real function
// instead of seven inlined copies per shading kernel (ncu r02a: the shading and camera kernels stall on instruction fetch as
// much as on memory -- 21 k and 7.5 k instructions, most of them inlined sampler and libm bodies).  The function reads the
// sampler parameters from constant memory (c_sampler, set at every render call) and the per-dimension hashes Hash(dimension,
// seed) from a constant table filled on the host with the same function.
#ifndef BPT_OUTLINE_SAMPLER
#define BPT_OUTLINE_SAMPLER 1
#endif
struct SamplerParams;
#if defined(__CUDACC__) && defined(BPT_SOBOL_DEVICE_TABLE)
static __device__ __noinline__ float2 ZSobolOutlined(uint64_t mortonIndex, int pixel, int dimension, int nd);
// The table-driven samplers other than ZSobol (PaddedSobol, Sobol, Halton) as one real function as well, so that kernels rendering
// with the default sampler do not carry three more inlined samplers at each of their seven call sites.  nd = 1 / 2: Get1D / Get2D
// at `dimension`; nd = 0: GetPixel2D.
static __device__ __noinline__ float2 QmcOutlined(int px, int py, int sampleIndex, int dimension, int nd);
// The per-pixel-sample set-up of the Sobol' / Halton samplers (SobolIntervalToIndex, the Halton index with its 64-bit divisions) runs
// INSIDE that function, per call: inlined into StartPixelSample it grew k_shade<Diffuse> from 22.9 k to 32.5 k instructions and the
// camera kernel from 7.1 k to 13.4 k (r02u: -11 % on the ZSobol benchmark, whose kernels stall on instruction fetch), and carrying the
// 64-bit index through the kernels cost them registers.  Those samplers pay for it; the default one does not.
#define BPT_QMC_OUTLINED 1
#else
#define BPT_QMC_OUTLINED 0
#endif
constexpr int kDimHashTable = 128;   // Hash(dimension, seed) for dimension < kDimHashTable

// One sampler object per path vertex evaluation: stateless function of (pixel, sampleIndex, dimension).
struct Sampler {
    const SamplerParams *sp;
    uint64_t mortonIndex;
    int dimension;
    int pixel;   // index into SamplerParams::prefix, -1: no table / pixel outside the film
    int px, py, sampleIndex;   // stratified / paddedsobol
    RNG rng;   // independent and stratified samplers
    int64_t sobolIndex;   // SobolSampler: the Sobol' index; HaltonSampler: haltonIndex
    BPT_HD void StartPixelSample(const SamplerParams *params, int pxIn, int pyIn, int index, int dim) {
        sp = params;
        dimension = dim;
        pixel = -1;
        px = pxIn; py = pyIn; sampleIndex = index;
#if BPT_QMC_OUTLINED && BPT_OUTLINE_SAMPLER && defined(__CUDA_ARCH__)
        // device: every sampler is evaluated as a stateless function of (pixel, sample index, dimension) by one of two real functions
        // (ZSobolOutlined / QmcOutlined); the RNG of the independent / stratified samplers is re-seeded and advanced to
        // index * 65536 + dimension there (its position is that at every call, samplers.h:457-460, 586-592), the global index of
        // Sobol' / Halton recomputed.  Only the default sampler keeps per-path state (Morton index, prefix-table slot).
        if (sp->type >= 4) dimension = dim > 2 ? dim : 2;
        if (sp->type == 0) {
            mortonIndex = (EncodeMorton2((uint32_t)pxIn, (uint32_t)pyIn) << sp->log2SamplesPerPixel) | (uint64_t)index;
            if (sp->prefix && pxIn >= 0 && pyIn >= 0 && pxIn < sp->prefixW && pyIn < sp->prefixH) pixel = pyIn * sp->prefixW + pxIn;
        }
        return;
#endif
        if (sp->type >= 4) { StartQmc(pxIn, pyIn, index, dim); return; }
        if (sp->type == 0) {
            mortonIndex = (EncodeMorton2((uint32_t)pxIn, (uint32_t)pyIn) << sp->log2SamplesPerPixel) | (uint64_t)index;
            if (sp->prefix && pxIn >= 0 && pyIn >= 0 && pxIn < sp->prefixW && pyIn < sp->prefixH) pixel = pyIn * sp->prefixW + pxIn;
        } else if (sp->type != 3) {   // (PaddedSobolSampler::StartPixelSample, samplers.h:389-393, keeps no other state)
            // IndependentSampler / StratifiedSampler::StartPixelSample (samplers.h:457-460, 586-592): Hash(Point2i p, int seed)
            uint64_t w[2] = {(uint64_t)(uint32_t)pxIn | ((uint64_t)(uint32_t)pyIn << 32), (uint64_t)(uint32_t)sp->seed};
            uint64_t h = MurmurHash64A(w, 12, 0);
            rng.SetSequence(h, MixBits(h));
            rng.Advance((int64_t)((uint64_t)index * 65536ull + (uint64_t)dim));
        }
    }
    // the samplers with a global sample index (types 4, 5) and pmj02bn (6)
    BPT_HD void StartQmc(int pxIn, int pyIn, int index, int dim) {
        if (sp->type == 6) { dimension = dim > 2 ? dim : 2; return; }   // PMJ02BNSampler::StartPixelSample (samplers.h:381-385)
        if (sp->type == 5) {   // HaltonSampler::StartPixelSample (samplers.h:53-71)
            int64_t haltonIndex = 0;
            int sampleStride = sp->baseScales[0] * sp->baseScales[1];
            if (sampleStride > 1) {
                auto mod = [](int a, int b) { int r = a - (a / b) * b; return r < 0 ? r + b : r; };
                const int pm[2] = {mod(pxIn, 128), mod(pyIn, 128)};
                for (int i = 0; i < 2; ++i) {
                    // InverseRadicalInverse(pm[i], base, baseExponents[i]) (lowdiscrepancy.h:104-113)
                    const int base = i == 0 ? 2 : 3;
                    uint64_t inverse = (uint64_t)pm[i], dimOffset = 0;
                    for (int k = 0; k < sp->baseExponents[i]; ++k) { uint64_t digit = inverse % base; inverse /= base; dimOffset = dimOffset * base + digit; }
                    haltonIndex += dimOffset * (sampleStride / sp->baseScales[i]) * sp->multInverse[i];
                }
                haltonIndex %= sampleStride;
            }
            haltonIndex += index * sampleStride;
            sobolIndex = haltonIndex;
            dimension = dim > 2 ? dim : 2;
            return;
        }
        if (sp->type == 4) {   // SobolSampler::StartPixelSample (samplers.h:502-506) + SobolIntervalToIndex (lowdiscrepancy.h:266-287)
            dimension = dim > 2 ? dim : 2;
            const uint32_t m = (uint32_t)sp->log2Scale;
            uint64_t frame = (uint64_t)index;
            if (m == 0) { sobolIndex = (int64_t)frame; return; }
            uint64_t idx = frame << (m << 1), delta = 0;
            for (int c = 0; frame; frame >>= 1, ++c) if (frame & 1) delta ^= sp->vdc[(m - 1) * 52 + c];
            uint64_t b = ((((uint64_t)(uint32_t)pxIn) << m) | (uint32_t)pyIn) ^ delta;
            for (int c = 0; b; b >>= 1, ++c) if (b & 1) idx ^= sp->vdcInv[(m - 1) * 52 + c];
            sobolIndex = (int64_t)idx;
            return;
        }
    }
    BPT_HD uint64_t GetSampleIndex() const {
        uint64_t sampleIndex = 0;
        bool pow2Samples = sp->log2SamplesPerPixel & 1;
        int lastDigit = pow2Samples ? 1 : 0;
        int top = sp->nBase4Digits - 1;
#if defined(__CUDA_ARCH__)
        if (pixel >= 0 && dimension < sp->prefixDims) {
            sampleIndex = __ldg(sp->prefix + (size_t)dimension * ((size_t)sp->prefixW * sp->prefixH) + pixel);
            top = sp->nLowDigits - 1;
        }
#endif
        return FinishSampleIndex(sampleIndex, top, lastDigit, pow2Samples);
    }
    // digits [lastDigit, top] of the permuted sample index (ZSobolSampler::GetSampleIndex, samplers.h:300-356), ORed into `sampleIndex`
    BPT_HD uint64_t FinishSampleIndex(uint64_t sampleIndex, int top, int lastDigit, bool pow2Samples, bool tail = true) const {
        // the 24 permutations of {0,1,2,3} (lowdiscrepancy.h:354-382), packed 2 bits per entry and keyed by the digit:
        // kPerm[d] >> 2p & 3 == permutations[p][d]
        const uint64_t kPerm0 = 0xfffaaa555000ull, kPerm1 = 0x0a5f05fa0fa5ull, kPerm2 = 0x6124dc2ce6deull, kPerm3 = 0x94817383b97bull;
        for (int i = top; i >= lastDigit; --i) {
            int digitShift = 2 * i - (pow2Samples ? 1 : 0);
            int digit = (int)((mortonIndex >> digitShift) & 3);
            uint64_t higherDigits = mortonIndex >> (digitShift + 2);
            uint64_t mix = MixBits(higherDigits ^ (uint64_t)(0x55555555u * (uint32_t)dimension)) >> 24;
            // mix % 24 in 32-bit arithmetic: mix < 2^40, 2^32 % 24 == 16
            uint32_t p = (((uint32_t)(mix >> 32) % 24u) * 16u + (uint32_t)mix % 24u) % 24u;
            uint64_t pk = digit == 0 ? kPerm0 : digit == 1 ? kPerm1 : digit == 2 ? kPerm2 : kPerm3;
            digit = (int)((pk >> (2 * p)) & 3);
            sampleIndex |= (uint64_t)digit << digitShift;
        }
        if (pow2Samples && tail) {
            int digit = (int)(mortonIndex & 1);
            sampleIndex |= (uint64_t)(digit ^ (int)(MixBits((mortonIndex >> 1) ^ (uint64_t)(0x55555555u * (uint32_t)dimension)) & 1));
        }
        return sampleIndex;
    }
    // StratifiedSampler::Get1D / Get2D (samplers.h:595-618), PaddedSobolSampler::Get1D / Get2D (samplers.h:395-421)
    BPT_HD float OtherGet1D() {
        uint64_t hash = Hash_pixel_dim_seed(px, py, dimension, sp->seed);
        if (sp->type == 2) {
            int n = sp->xSamples * sp->ySamples;
            int stratum = PermutationElement((uint32_t)sampleIndex, (uint32_t)n, (uint32_t)hash);
            ++dimension;
            float delta = sp->jitter ? rng.UniformFloat() : 0.5f;
            return (stratum + delta) / n;
        }
        int index = PermutationElement((uint32_t)sampleIndex, (uint32_t)sp->samplesPerPixel, (uint32_t)hash);
        ++dimension;
        return SobolSampleRand((uint64_t)index, 0, sp->randomize, (uint32_t)(hash >> 32));
    }
    BPT_HD Vec2 OtherGet2D() {
        uint64_t hash = Hash_pixel_dim_seed(px, py, dimension, sp->seed);
        if (sp->type == 2) {
            int n = sp->xSamples * sp->ySamples;
            int stratum = PermutationElement((uint32_t)sampleIndex, (uint32_t)n, (uint32_t)hash);
            dimension += 2;
            int x = stratum % sp->xSamples, y = stratum / sp->xSamples;
            float dx = sp->jitter ? rng.UniformFloat() : 0.5f;
            float dy = sp->jitter ? rng.UniformFloat() : 0.5f;
            return {(x + dx) / sp->xSamples, (y + dy) / sp->ySamples};
        }
        int index = PermutationElement((uint32_t)sampleIndex, (uint32_t)sp->samplesPerPixel, (uint32_t)hash);
        dimension += 2;
        return {SobolSampleRand((uint64_t)index, 0, sp->randomize, (uint32_t)hash), SobolSampleRand((uint64_t)index, 1, sp->randomize, (uint32_t)(hash >> 32))};
    }
    // SobolSampler::SampleDimension (samplers.h:544-557); randomize < 0: NoRandomizer
    BPT_HD float SobolDimension(int dim, int randomize) const {
        const uint32_t *C = sp->sobolFull + (size_t)dim * 52;
        uint32_t v = 0;
        uint64_t a = (uint64_t)sobolIndex;
        for (int i = 0; a != 0; a >>= 1, i++) if (a & 1) v ^= C[i];
        if (randomize > 0) {
            uint32_t hash = (uint32_t)Hash_i32x2(dim, sp->seed);
            if (randomize == RANDOMIZE_PERMUTE_DIGITS) v = hash ^ v;
            else if (randomize == RANDOMIZE_FAST_OWEN) v = FastOwenScramble(v, hash);
            else v = OwenScramble(v, hash);
        }
        return stdmin((float)v * 0x1p-32f, kOneMinusEpsilon);
    }
    // RadicalInverse / ScrambledRadicalInverse / OwenScrambledRadicalInverse (lowdiscrepancy.h:86-158) of the sample's Halton index
    BPT_HD static float RadicalInverse(unsigned base, uint64_t a) {
        uint64_t limit = ~0ull / base - base;
        float invBase = (float)1 / (float)base, invBaseM = 1;
        uint64_t reversedDigits = 0;
        while (a && reversedDigits < limit) {
            uint64_t next = a / base;
            uint64_t digit = a - next * base;
            reversedDigits = reversedDigits * base + digit;
            invBaseM *= invBase;
            a = next;
        }
        return stdmin(reversedDigits * invBaseM, kOneMinusEpsilon);
    }
    BPT_HD float HaltonDimension(int dim) const {
        const unsigned base = (unsigned)sp->primes[dim];
        uint64_t a = (uint64_t)sobolIndex;
        if (sp->randomize == RANDOMIZE_NONE) return RadicalInverse(base, a);
        uint64_t limit = ~0ull / base - base;
        float invBase = (float)1 / (float)base, invBaseM = 1;
        uint64_t reversedDigits = 0;
        int digitIndex = 0;
        if (sp->randomize == RANDOMIZE_PERMUTE_DIGITS) {
            const uint16_t *perm = sp->haltonPerms + sp->haltonPermOffset[dim];
            while (1 - (base - 1) * invBaseM < 1 && reversedDigits < limit) {
                uint64_t next = a / base;
                int digitValue = (int)(a - next * base);
                reversedDigits = reversedDigits * base + perm[digitIndex * (int)base + digitValue];
                invBaseM *= invBase;
                ++digitIndex;
                a = next;
            }
            return stdmin(invBaseM * reversedDigits, kOneMinusEpsilon);
        }
        const uint32_t hash = (uint32_t)MixBits((uint64_t)(1 + (dim << 4)));
        while (1 - invBaseM < 1 && reversedDigits < limit) {
            uint64_t next = a / base;
            int digitValue = (int)(a - next * base);
            uint32_t digitHash = (uint32_t)MixBits(hash ^ reversedDigits);
            digitValue = PermutationElement((uint32_t)digitValue, base, digitHash);
            reversedDigits = reversedDigits * base + digitValue;
            invBaseM *= invBase;
            ++digitIndex;
            a = next;
        }
        return stdmin(invBaseM * reversedDigits, kOneMinusEpsilon);
    }
    // PMJ02BNSampler::Get1D / Get2D / GetPixel2D (samplers.h:388-427), BlueNoise (util/bluenoise.h:25-30), GetPMJ02BNSample
    // (util/pmj02tables.h:26-41)
    BPT_HD float BlueNoise(int textureIndex) const {
        textureIndex %= 48;
        int x = px % 128, y = py % 128;
        return sp->blueNoise[((size_t)textureIndex * 128 + x) * 128 + y] / 65535.f;
    }
    BPT_HD float Pmj1D() {
        uint64_t hash = Hash_pixel_dim_seed(px, py, dimension, sp->seed);
        int index = PermutationElement((uint32_t)sampleIndex, (uint32_t)sp->samplesPerPixel, (uint32_t)hash);
        float delta = BlueNoise(dimension);
        ++dimension;
        return stdmin((index + delta) / sp->samplesPerPixel, kOneMinusEpsilon);
    }
    BPT_HD Vec2 Pmj2D() {
        int index = sampleIndex;
        int pmjInstance = dimension / 2;
        if (pmjInstance >= 5) {
            uint64_t hash = Hash_pixel_dim_seed(px, py, dimension, sp->seed);
            index = PermutationElement((uint32_t)sampleIndex, (uint32_t)sp->samplesPerPixel, (uint32_t)hash);
        }
        const uint32_t *smp = sp->pmjSamples + ((size_t)(pmjInstance % 5) * 65536 + (size_t)(index % 65536)) * 2;
        Vec2 u((float)smp[0] * 0x1p-32f, (float)smp[1] * 0x1p-32f);
        u.x += BlueNoise(dimension); u.y += BlueNoise(dimension + 1);
        if (u.x >= 1) u.x -= 1;
        if (u.y >= 1) u.y -= 1;
        dimension += 2;
        return {stdmin(u.x, kOneMinusEpsilon), stdmin(u.y, kOneMinusEpsilon)};
    }
    // Sampler::GetPixel2D: the film-plane sample of GetCameraSample (samplers.h:797-814); SobolSampler (:525-536), HaltonSampler
    // (:90-93) and PMJ02BNSampler (:399-403) have their own
    BPT_HD Vec2 PmjPixel2D() const {
        int tx = px % sp->pixelTileSize, ty = py % sp->pixelTileSize;
        int offset = (tx + ty * sp->pixelTileSize) * sp->samplesPerPixel;
        const float *q = sp->pmjPixelSamples + 2 * (size_t)(offset + sampleIndex);
        return {q[0], q[1]};
    }
    BPT_HD Vec2 GetPixel2D() {
#if BPT_QMC_OUTLINED && BPT_OUTLINE_SAMPLER && defined(__CUDA_ARCH__)
        if (sp->type >= 4) { float2 r = QmcOutlined(px, py, sampleIndex, dimension, 0); return {r.x, r.y}; }
        return Get2D();
#else
        if (sp->type == 6) return PmjPixel2D();
        if (sp->type == 5)
            return {RadicalInverse(2, (uint64_t)sobolIndex >> sp->baseExponents[0]), RadicalInverse(3, (uint64_t)(sobolIndex / sp->baseScales[1]))};
        if (sp->type != 4) return Get2D();
        float u[2] = {SobolDimension(0, 0), SobolDimension(1, 0)};
        const int pix[2] = {px, py};
        for (int d = 0; d < 2; ++d) u[d] = Clampf(u[d] * sp->scale - pix[d], 0.f, kOneMinusEpsilon);
        return {u[0], u[1]};
#endif
    }
    BPT_HD float Get1D() {
#if BPT_QMC_OUTLINED && BPT_OUTLINE_SAMPLER && defined(__CUDA_ARCH__)
        if (sp->type == 4 && dimension >= 1024) dimension = 2;
        if (sp->type == 5 && dimension >= 1000) dimension = 2;
        float2 r = sp->type == 0 ? ZSobolOutlined(mortonIndex, pixel, dimension, 1) : QmcOutlined(px, py, sampleIndex, dimension, 1);
        ++dimension;
        return r.x;
#else
        if (sp->type == 6) return Pmj1D();
        if (sp->type == 5) {
            if (dimension >= 1000) dimension = 2;
            return HaltonDimension(dimension++);
        }
        if (sp->type == 4) {
            if (dimension >= 1024) dimension = 2;
            return SobolDimension(dimension++, sp->randomize);
        }
        if (sp->type == 1) { ++dimension; return rng.UniformFloat(); }
        if (sp->type >= 2) return OtherGet1D();
        uint64_t sampleIndex = GetSampleIndex();
        ++dimension;
        uint32_t sampleHash = (uint32_t)Hash_i32x2(dimension, sp->seed);
        return SobolSampleRand(sampleIndex, 0, sp->randomize, sampleHash);
#endif
    }
    BPT_HD Vec2 Get2D() {
#if BPT_QMC_OUTLINED && BPT_OUTLINE_SAMPLER && defined(__CUDA_ARCH__)
        if (sp->type == 4 && dimension + 1 >= 1024) dimension = 2;
        if (sp->type == 5 && dimension + 1 >= 1000) dimension = 2;
        float2 r = sp->type == 0 ? ZSobolOutlined(mortonIndex, pixel, dimension, 2) : QmcOutlined(px, py, sampleIndex, dimension, 2);
        dimension += 2;
        return {r.x, r.y};
#else
        if (sp->type == 6) return Pmj2D();
        if (sp->type == 5) {
            if (dimension + 1 >= 1000) dimension = 2;
            int dim = dimension;
            dimension += 2;
            return {HaltonDimension(dim), HaltonDimension(dim + 1)};
        }
        if (sp->type == 4) {
            if (dimension + 1 >= 1024) dimension = 2;
            Vec2 u(SobolDimension(dimension, sp->randomize), SobolDimension(dimension + 1, sp->randomize));
            dimension += 2;
            return u;
        }
        if (sp->type == 1) { dimension += 2; float a = rng.UniformFloat(); float b = rng.UniformFloat(); return {a, b}; }
        if (sp->type >= 2) return OtherGet2D();
        uint64_t sampleIndex = GetSampleIndex();
        dimension += 2;
        uint64_t bits = Hash_i32x2(dimension, sp->seed);
        return {SobolSampleRand(sampleIndex, 0, sp->randomize, (uint32_t)bits),
                SobolSampleRand(sampleIndex, 1, sp->randomize, (uint32_t)(bits >> 32))};
#endif
    }
};

#if defined(__CUDACC__) && defined(BPT_SOBOL_DEVICE_TABLE)
__constant__ SamplerParams c_sampler;
__constant__ uint64_t c_dimHash[kDimHashTable];
static __device__ __noinline__ float2 QmcOutlined(int px, int py, int sampleIndex, int dimension, int nd) {
    Sampler s;
    s.sp = &c_sampler; s.px = px; s.py = py; s.sampleIndex = sampleIndex; s.sobolIndex = 0; s.pixel = -1;
    if (c_sampler.type >= 4) s.StartQmc(px, py, sampleIndex, 0);   // Sobol' / Halton global index
    const long long sobolIndex = s.sobolIndex;
    s.dimension = dimension;
    float2 r = make_float2(0.f, 0.f);
    const int type = c_sampler.type;
    if (nd == 0) {   // GetPixel2D of the Sobol', Halton and pmj02bn samplers (the member functions' own code, not the dispatch above)
        if (type == 6) { Vec2 u = s.PmjPixel2D(); r.x = u.x; r.y = u.y; }
        else if (type == 5) { r.x = Sampler::RadicalInverse(2, (uint64_t)sobolIndex >> c_sampler.baseExponents[0]); r.y = Sampler::RadicalInverse(3, (uint64_t)(sobolIndex / c_sampler.baseScales[1])); }
        else {
            float u[2] = {s.SobolDimension(0, 0), s.SobolDimension(1, 0)};
            const int pix[2] = {px, py};
            for (int d = 0; d < 2; ++d) u[d] = Clampf(u[d] * c_sampler.scale - pix[d], 0.f, kOneMinusEpsilon);
            r.x = u[0]; r.y = u[1];
        }
        return r;
    }
    if (type <= 3) {
        if (type != 3) {   // IndependentSampler / StratifiedSampler: the RNG as StartPixelSample + the Get calls so far have left it
            uint64_t w[2] = {(uint64_t)(uint32_t)px | ((uint64_t)(uint32_t)py << 32), (uint64_t)(uint32_t)c_sampler.seed};
            uint64_t h = MurmurHash64A(w, 12, 0);
            s.rng.SetSequence(h, MixBits(h));
            s.rng.Advance((int64_t)((uint64_t)sampleIndex * 65536ull + (uint64_t)dimension));
        }
        if (type == 1) { r.x = s.rng.UniformFloat(); if (nd == 2) r.y = s.rng.UniformFloat(); }
        else if (nd == 1) r.x = s.OtherGet1D(); else { Vec2 u = s.OtherGet2D(); r.x = u.x; r.y = u.y; }
    } else if (type == 6) {
        if (nd == 1) r.x = s.Pmj1D(); else { Vec2 u = s.Pmj2D(); r.x = u.x; r.y = u.y; }
    } else if (type == 4) {
        r.x = s.SobolDimension(dimension, c_sampler.randomize);
        if (nd == 2) r.y = s.SobolDimension(dimension + 1, c_sampler.randomize);
    } else {
        r.x = s.HaltonDimension(dimension);
        if (nd == 2) r.y = s.HaltonDimension(dimension + 1);
    }
    return r;
}
static __device__ __noinline__ float2 ZSobolOutlined(uint64_t mortonIndex, int pixel, int dimension, int nd) {
    Sampler s;
    s.sp = &c_sampler; s.mortonIndex = mortonIndex; s.pixel = pixel; s.dimension = dimension;
    uint64_t sampleIndex = s.GetSampleIndex();
    int dim = dimension + nd;
    uint64_t bits = dim < kDimHashTable ? c_dimHash[dim] : Hash_i32x2(dim, c_sampler.seed);
    float2 r;
    r.x = SobolSampleRand(sampleIndex, 0, c_sampler.randomize, (uint32_t)bits);
    r.y = nd == 2 ? SobolSampleRand(sampleIndex, 1, c_sampler.randomize, (uint32_t)(bits >> 32)) : 0.f;
    return r;
}
#endif

// ------------------------------------------------------------------ sampling routines
BPT_HD float SampleVisibleWavelengths(float u) { return 538 - 138.888889f * ATanhF(0.85691062f - 1.82750197f * u); }
BPT_HD float VisibleWavelengthsPDF(float lambda) {
    if (lambda < 360 || lambda > 830) return 0;
    return 0.0039398042f / Sqr(CoshF(0.0072f * (lambda - 538)));
}

BPT_HD Vec2 SampleUniformDiskConcentric(Vec2 u) {
    Vec2 uOffset(2 * u.x - 1, 2 * u.y - 1);
    if (uOffset.x == 0 && uOffset.y == 0) return {0, 0};
    float theta, r;
    if (fabsf(uOffset.x) > fabsf(uOffset.y)) {
        r = uOffset.x;
        theta = kPiOver4 * (uOffset.y / uOffset.x);
    } else {
        r = uOffset.y;
        theta = kPiOver2 - kPiOver4 * (uOffset.x / uOffset.y);
    }
    return {r * CosF(theta), r * SinF(theta)};
}
BPT_HD Vec3 SampleCosineHemisphere(Vec2 u) {
    Vec2 d = SampleUniformDiskConcentric(u);
    float z = SafeSqrt(1 - Sqr(d.x) - Sqr(d.y));
    return {d.x, d.y, z};
}
BPT_HD float CosineHemispherePDF(float cosTheta) { return cosTheta * kInvPi; }

BPT_HD void SampleUniformTriangle(Vec2 u, float b[3]) {
    float b0, b1;
    if (u.x < u.y) { b0 = u.x / 2; b1 = u.y - b0; }
    else { b1 = u.y / 2; b0 = u.x - b1; }
    b[0] = b0; b[1] = b1; b[2] = 1 - b0 - b1;
}
BPT_HD float SampleLinear(float u, float a, float b) {
    if (u == 0 && a == 0) return 0;
    float x = u * (a + b) / (a + sqrtf(Lerp(u, Sqr(a), Sqr(b))));
    return stdmin(x, kOneMinusEpsilon);
}
// SampleTent (util/sampling.h:196-201) with SampleDiscrete({0.5, 0.5}, u, nullptr, &u) (util/sampling.h:79-113) written out
BPT_HD float SampleTent(float u, float r) {
    const float sumWeights = 0.5f + 0.5f;
    float up = u * sumWeights;
    if (up == sumWeights) up = NextFloatDown(up);
    int offset = 0;
    float sum = 0;
    while (sum + 0.5f <= up) { sum += 0.5f; ++offset; }
    u = stdmin((up - sum) / 0.5f, kOneMinusEpsilon);
    if (offset == 0) return -r + r * SampleLinear(u, 0, 1);
    return r * SampleLinear(u, 1, 0);
}
BPT_HD float BilinearPDF(Vec2 p, const float w[4]) {
    if (p.x < 0 || p.x > 1 || p.y < 0 || p.y > 1) return 0;
    if (w[0] + w[1] + w[2] + w[3] == 0) return 1;
    return 4 * ((1 - p.x) * (1 - p.y) * w[0] + p.x * (1 - p.y) * w[1] + (1 - p.x) * p.y * w[2] + p.x * p.y * w[3]) /
           (w[0] + w[1] + w[2] + w[3]);
}
BPT_HD Vec2 SampleBilinear(Vec2 u, const float w[4]) {
    Vec2 p;
    p.y = SampleLinear(u.y, w[0] + w[1], w[2] + w[3]);
    p.x = SampleLinear(u.x, Lerp(p.y, w[0], w[2]), Lerp(p.y, w[1], w[3]));
    return p;
}
BPT_HD float SphericalTriangleArea(Vec3 a, Vec3 b, Vec3 c) {
    return fabsf(2 * ATan2F(Dot(a, Cross(b, c)), 1 + Dot(a, b) + Dot(a, c) + Dot(b, c)));
}

// util/sampling.cpp SampleSphericalTriangle.  Returns false for the "{}" (all-zero barycentrics) exits.
BPT_HD void SampleSphericalTriangle(Pt3 v0, Pt3 v1, Pt3 v2, Pt3 p, Vec2 u, float *pdf, float bary[3]) {
    *pdf = 0;
    bary[0] = bary[1] = bary[2] = 0;
    Vec3 a = v0 - p, b = v1 - p, c = v2 - p;
    a = Normalize(a); b = Normalize(b); c = Normalize(c);
    Vec3 n_ab = Cross(a, b), n_bc = Cross(b, c), n_ca = Cross(c, a);
    if (LengthSquared(n_ab) == 0 || LengthSquared(n_bc) == 0 || LengthSquared(n_ca) == 0) return;
    n_ab = Normalize(n_ab); n_bc = Normalize(n_bc); n_ca = Normalize(n_ca);
    float alpha = AngleBetween(n_ab, -n_ca);
    float beta = AngleBetween(n_bc, -n_ab);
    float gamma = AngleBetween(n_ca, -n_bc);
    float A_pi = alpha + beta + gamma;
    float Ap_pi = Lerp(u.x, kPi, A_pi);
    float A = A_pi - kPi;
    *pdf = (A <= 0) ? 0 : 1 / A;
    float cosAlpha = CosF(alpha), sinAlpha = SinF(alpha);
    float sinPhi = SinF(Ap_pi) * cosAlpha - CosF(Ap_pi) * sinAlpha;
    float cosPhi = CosF(Ap_pi) * cosAlpha + SinF(Ap_pi) * sinAlpha;
    float k1 = cosPhi + cosAlpha;
    float k2 = sinPhi - sinAlpha * Dot(a, b);
    float cosBp = (k2 + (DifferenceOfProducts(k2, cosPhi, k1, sinPhi)) * cosAlpha) /
                  ((SumOfProducts(k2, sinPhi, k1, cosPhi)) * sinAlpha);
    cosBp = Clampf(cosBp, -1, 1);
    float sinBp = SafeSqrt(1 - Sqr(cosBp));
    Vec3 cp = cosBp * a + sinBp * Normalize(GramSchmidt(c, a));
    float cosTheta = 1 - u.y * (1 - Dot(cp, b));
    float sinTheta = SafeSqrt(1 - Sqr(cosTheta));
    Vec3 w = cosTheta * b + sinTheta * Normalize(GramSchmidt(cp, b));
    Vec3 e1 = v1 - v0, e2 = v2 - v0;
    Vec3 s1 = Cross(w, e2);
    float divisor = Dot(s1, e1);
    if (divisor == 0) { bary[0] = bary[1] = bary[2] = 1.f / 3.f; return; }
    float invDivisor = 1 / divisor;
    Vec3 s = p - v0;
    float b1 = Dot(s, s1) * invDivisor;
    float b2 = Dot(w, Cross(s, e1)) * invDivisor;
    b1 = Clampf(b1, 0, 1);
    b2 = Clampf(b2, 0, 1);
    if (b1 + b2 > 1) {
        b1 /= b1 + b2;
        b2 /= b1 + b2;   // (sic) the reference divides by the already-updated b1 + b2
    }
    bary[0] = 1 - b1 - b2; bary[1] = b1; bary[2] = b2;
}

BPT_HD Vec2 InvertSphericalTriangleSample(Pt3 v0, Pt3 v1, Pt3 v2, Pt3 p, Vec3 w) {
    Vec3 a = v0 - p, b = v1 - p, c = v2 - p;
    a = Normalize(a); b = Normalize(b); c = Normalize(c);
    Vec3 n_ab = Cross(a, b), n_bc = Cross(b, c), n_ca = Cross(c, a);
    if (LengthSquared(n_ab) == 0 || LengthSquared(n_bc) == 0 || LengthSquared(n_ca) == 0) return {0, 0};
    n_ab = Normalize(n_ab); n_bc = Normalize(n_bc); n_ca = Normalize(n_ca);
    float alpha = AngleBetween(n_ab, -n_ca);
    float beta = AngleBetween(n_bc, -n_ab);
    float gamma = AngleBetween(n_ca, -n_bc);
    Vec3 cp = Normalize(Cross(Cross(b, w), Cross(c, a)));
    if (Dot(cp, a + c) < 0) cp = -cp;
    float u0;
    if (Dot(a, cp) > 0.99999847691f) u0 = 0;
    else {
        Vec3 n_cpb = Cross(cp, b), n_acp = Cross(a, cp);
        if (LengthSquared(n_cpb) == 0 || LengthSquared(n_acp) == 0) return {0.5f, 0.5f};
        n_cpb = Normalize(n_cpb); n_acp = Normalize(n_acp);
        float Ap = alpha + AngleBetween(n_ab, n_cpb) + AngleBetween(n_acp, -n_cpb) - kPi;
        float A = alpha + beta + gamma - kPi;
        u0 = Ap / A;
    }
    float u1 = (1 - Dot(w, b)) / (1 - Dot(cp, b));
    return {Clampf(u0, 0, 1), Clampf(u1, 0, 1)};
}

// FindInterval (util/math.h:508-519) over a float array with predicate arr[i] <= u
BPT_HD int FindIntervalLE(const float *arr, int sz, float u) {
    int size = sz - 2, first = 1;
    while (size > 0) {
        int half = size >> 1, middle = first + half;
        bool pr = arr[middle] <= u;
        first = pr ? middle + 1 : first;
        size = pr ? size - (half + 1) : half;
    }
    return Clampi(first - 1, 0, sz - 2);
}
BPT_HD int FindIntervalLT(const float *arr, int sz, float u) {
    int size = sz - 2, first = 1;
    while (size > 0) {
        int half = size >> 1, middle = first + half;
        bool pr = arr[middle] < u;
        first = pr ? middle + 1 : first;
        size = pr ? size - (half + 1) : half;
    }
    return Clampi(first - 1, 0, sz - 2);
}

// PiecewiseConstant2D stored flat: func[nv][nu], cdf[nv][nu+1], funcInt[nv], then marginal func[nv], cdf[nv+1].
struct PC2DView {
    const float *func;       // nv*nu (already |.|)
    const float *cdf;        // nv*(nu+1)
    const float *funcInt;    // nv
    const float *mfunc;      // nv
    const float *mcdf;       // nv+1
    float mfuncInt;
    int nu, nv;
    float min0, max0, min1, max1;
};
BPT_HD float PC1DSample(const float *func, const float *cdf, int n, float funcInt, float mn, float mx, float u,
                        float *pdf, int *offset) {
    int o = FindIntervalLE(cdf, n + 1, u);
    *offset = o;
    float du = u - cdf[o];
    if (cdf[o + 1] - cdf[o] > 0) du /= cdf[o + 1] - cdf[o];
    *pdf = (funcInt > 0) ? func[o] / funcInt : 0;
    return Lerp((o + du) / n, mn, mx);
}
BPT_HD Vec2 PC2DSample(const PC2DView &d, Vec2 u, float *pdf, int *ou, int *ov) {
    float pdf0, pdf1;
    float d1 = PC1DSample(d.mfunc, d.mcdf, d.nv, d.mfuncInt, d.min1, d.max1, u.y, &pdf1, ov);
    float d0 = PC1DSample(d.func + (size_t)(*ov) * d.nu, d.cdf + (size_t)(*ov) * (d.nu + 1), d.nu, d.funcInt[*ov],
                          d.min0, d.max0, u.x, &pdf0, ou);
    *pdf = pdf0 * pdf1;
    return {d0, d1};
}
BPT_HD float PC2DPDF(const PC2DView &d, Vec2 pr) {
    // domain.Offset(pr)
    float ox = pr.x - d.min0, oy = pr.y - d.min1;
    if (d.max0 > d.min0) ox /= d.max0 - d.min0;
    if (d.max1 > d.min1) oy /= d.max1 - d.min1;
    int iu = Clampi((int)(ox * d.nu), 0, d.nu - 1);
    int iv = Clampi((int)(oy * d.nv), 0, d.nv - 1);
    return d.func[(size_t)iv * d.nu + iu] / d.mfuncInt;
}

// ---- host: the tables and constants of the Halton sampler (HaltonSampler::HaltonSampler samplers.cpp:32-52,
// ComputeRadicalInversePermutations lowdiscrepancy.cpp:47-55, DigitPermutation lowdiscrepancy.h:30-49)
inline std::vector<int> FirstPrimes(int n) {
    std::vector<int> primes;
    for (int k = 2; (int)primes.size() < n; ++k) {
        bool pr = true;
        for (int q : primes) { if (q * q > k) break; if (k % q == 0) { pr = false; break; } }
        if (pr) primes.push_back(k);
    }
    return primes;
}
inline void BuildHaltonPermutations(const std::vector<int> &primes, int seed, std::vector<uint16_t> *perms, std::vector<uint32_t> *offsets) {
    perms->clear();
    offsets->assign(primes.size(), 0);
    for (size_t i = 0; i < primes.size(); ++i) {
        const int base = primes[i];
        int nDigits = 0;
        float invBase = (float)1 / (float)base, invBaseM = 1;
        while (1 - (base - 1) * invBaseM < 1) { ++nDigits; invBaseM *= invBase; }
        (*offsets)[i] = (uint32_t)perms->size();
        for (int digitIndex = 0; digitIndex < nDigits; ++digitIndex) {
            // Hash(base, digitIndex, seed): the three 32-bit arguments packed back to back (util/hash.h:99-107)
            uint64_t w[2] = {(uint64_t)(uint32_t)base | ((uint64_t)(uint32_t)digitIndex << 32), (uint64_t)(uint32_t)seed};
            uint64_t dseed = MurmurHash64A(w, 12, 0);
            for (int digitValue = 0; digitValue < base; ++digitValue) perms->push_back((uint16_t)PermutationElement((uint32_t)digitValue, (uint32_t)base, (uint32_t)dseed));
        }
    }
}
// PMJ02BNSampler::PMJ02BNSampler (samplers.cpp:173-207): the points of the first pmj02bn set sorted into the pixels of a
// pixelTileSize x pixelTileSize tile, samplesPerPixel each (GetPMJ02BNSample's CPU branch converts the fixed-point values in double)
inline bool BuildPmjPixelSamples(const uint32_t *pmjSamples, int samplesPerPixel, int *pixelTileSize, std::vector<float> *pixelSamples) {
    auto log4 = [](int v) { int l = 0; while (v >>= 2) ++l; return l; };
    auto roundUpPow4 = [](int v) { int r = 1; while (r < v) r <<= 2; return r; };
    if (samplesPerPixel > 65536 || samplesPerPixel < 1) return false;
    *pixelTileSize = 1 << (log4(65536) - log4(roundUpPow4(samplesPerPixel)));
    const int T = *pixelTileSize;
    pixelSamples->assign((size_t)2 * T * T * samplesPerPixel, 0.f);
    std::vector<int> nStored((size_t)T * T, 0);
    for (int i = 0; i < 65536; ++i) {
        float px = (float)(pmjSamples[2 * i] * 0x1p-32), py = (float)(pmjSamples[2 * i + 1] * 0x1p-32);
        px *= T; py *= T;
        int pixelOffset = int(px) + int(py) * T;
        if (nStored[pixelOffset] == samplesPerPixel) continue;
        int sampleOffset = pixelOffset * samplesPerPixel + nStored[pixelOffset];
        (*pixelSamples)[2 * (size_t)sampleOffset] = px - std::floor(px);
        (*pixelSamples)[2 * (size_t)sampleOffset + 1] = py - std::floor(py);
        ++nStored[pixelOffset];
    }
    for (int c : nStored) if (c != samplesPerPixel) return false;
    return true;
}
inline void HaltonPixelConstants(int xres, int yres, SamplerParams *sp) {
    const int fullRes[2] = {xres, yres};
    for (int i = 0; i < 2; ++i) {
        int base = (i == 0) ? 2 : 3, scale = 1, exp = 0;
        while (scale < (fullRes[i] < 128 ? fullRes[i] : 128)) { scale *= base; ++exp; }
        sp->baseScales[i] = scale; sp->baseExponents[i] = exp;
    }
    struct G { static void ext(uint64_t a, uint64_t b, int64_t *x, int64_t *y) { if (b == 0) { *x = 1; *y = 0; return; } int64_t d = (int64_t)(a / b), xp, yp; ext(b, a % b, &xp, &yp); *x = yp; *y = xp - (d * yp); } };
    auto multiplicativeInverse = [](int64_t a, int64_t n) { int64_t x, y; G::ext((uint64_t)a, (uint64_t)n, &x, &y); int64_t r = x - (x / n) * n; return (int)(r < 0 ? r + n : r); };
    sp->multInverse[0] = multiplicativeInverse(sp->baseScales[1], sp->baseScales[0]);
    sp->multInverse[1] = multiplicativeInverse(sp->baseScales[0], sp->baseScales[1]);
}

}  // namespace bpt
