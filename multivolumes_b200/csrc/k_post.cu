// k_post.cu — temporal anti-aliasing and tone map as one streaming image kernel.
//
// Replaces CSTemporalAA (MultiVolumes/Content/Shaders/CSTemporalAA.hlsl:254-336, compiled with
// ALPHA_BOUND = 1.0, _USE_YCOCG_, _VARIANCE_AABB_; host ObjectRenderer.cpp:256-283) and PSToneMap
// (PSToneMap.hlsl:19-28; host ObjectRenderer.cpp:285-300). The tone map runs in the epilogue of the
// TAA on the value just computed (rounded through RGBA16F exactly as the intermediate render target
// of the reference would), so the frame is read once and both outputs are written once:
// 8 B colour + 8 B history + 4 B velocity in, 8 B + 4 B out per pixel; HBM-bound.
#include "mv_internal.h"

namespace mv {

namespace {

// Stated evaluation order of this pass (the test oracle states the same one): YCoCg as sums and doublings, every division as a
// multiplication by the correctly rounded reciprocal (rcp, mv_math.cuh) or by the reciprocal constant, fused multiply-adds
// (fma1) exactly where written. dxc compiles the reference with fast-math: its DXIL multiplies by reciprocals too.
MV_D V3 rgb_to_ycocg(V3 rgb)   // :78-85: (1 2 1; 2 0 -2; -1 2 -1)
{
    const float g2 = rgb.y + rgb.y;
    return {(rgb.x + g2) + rgb.z, (rgb.x - rgb.z) + (rgb.x - rgb.z), (g2 - rgb.x) - rgb.z};
}
MV_D V3 TM(V3 hdr) { const V3 c = rgb_to_ycocg(hdr); const float q = rcp(4.0f + c.x); return {c.x * q, c.y * q, c.z * q}; }   // :106-114
// :119-128 with :90-101 folded in: (c * (4 / (1 - c.x))) * 0.25 = c * rcp(1 - c.x) — scaling by 4 and by 0.25 is exact
MV_D V3 ITM(V3 c) { const float q = rcp(1.0f - c.x); const float y = c.x * q, co = c.y * q, cg = c.z * q; return {y + co - cg, y + cg, y - co - cg}; }
MV_D float lerpf(float a, float b, float t) { return fma1(b - a, t, a); }   // lerp as one fused multiply-add

// PSToneMap.hlsl:19-28 + the RGBA8_UNORM render-target write, for ONE channel value. The tone map's input is an RGBA16F
// texel, so the whole function is a table over the 65536 half patterns (k_build_tone_lut fills it with exactly this code).
MV_D unsigned char tone_map_channel(float v)
{
    v *= kToneScale / (v + kToneBias);
    v = pow125(fabsf(v));
    float sat = saturate(v);
    if (!(v == v)) sat = 0.0f;
    return (unsigned char)floorf(sat * 255.0f + 0.5f);
}

__global__ void __launch_bounds__(256) k_build_tone_lut(unsigned char* lut)
{
    const uint32_t h = blockIdx.x * 256 + threadIdx.x;
    lut[h] = tone_map_channel(f16_to_f32((uint16_t)h));
}

MV_D uchar4 tone_map(uint2 taaTexel, const unsigned char* __restrict__ lut)
{
    return make_uchar4(__ldg(lut + (taaTexel.x & 0xffffu)), __ldg(lut + (taaTexel.x >> 16)), __ldg(lut + (taaTexel.y & 0xffffu)), 255);
}

struct PostArgs {
    const uint2* color;       // current frame, RGBA16F
    const uint2* history;     // previous TAA output
    const uint32_t* velocity; // RG16F
    uint2* out;               // TAA output (next frame's history)
    uchar4* backBuffer;       // RGBA8
    uchar4* peerBackBuffer;   // rank 0's back buffer when this rank resolves a band of a multi-GPU frame
    uint2* peerOut[kMaxPeers]; // the peers' copies of `out` (multi-GPU, peers mapped): the next frame's history fetch of ANY rank may land on these rows
    int numPeers;
    float invW, invH;          // 1 / W, 1 / H (fp32 quotients, computed once on the host)
    int velocityGiven;         // 0: the velocity field is all zero (none was given): its five taps need not be read
    int peerAllRows;           // 0: only the first and last row of each stripe / band go to the peers (static velocity field: the history fetch stays within one row)
    int W, H, row0, row1;
    int stripeH, rank, world;   // stripeH > 0: interleaved stripes instead of the band
    int taaOn;
    const unsigned char* toneLut;   // [65536] tone map + RGBA8 quantisation of every half pattern
};

MV_D V4 load_c(const uint2* img, int x, int y, int W, int H)   // Texture2D[] load: out of bounds -> 0
{
    if (x < 0 || y < 0 || x >= W || y >= H) return {0.0f, 0.0f, 0.0f, 0.0f};
    return unpack_half4(__ldg(img + (size_t)y * W + x));
}
MV_D V2 load_v(const uint32_t* img, int x, int y, int W, int H)
{
    if (x < 0 || y < 0 || x >= W || y >= H) return {0.0f, 0.0f};
    const uint32_t p = __ldg(img + (size_t)y * W + x);
    return {f16_to_f32((uint16_t)(p & 0xffffu)), f16_to_f32((uint16_t)(p >> 16))};
}

constexpr int kPostW = 32, kPostH = 8;                    // pixels per CTA: a warp is 32 pixels of one row
constexpr int kTileW = kPostW + 2, kTileH = kPostH + 2;   // + one-pixel halo for the 3x3 neighbourhood

// The tone-mapped YCoCg colour TM(c) of every pixel is needed by its eight neighbours as well
// (NeighborMinMax); it is computed once per pixel into a shared-memory tile (3 IEEE divides each)
// instead of nine times, with the same expressions, so the result is unchanged.
__global__ void __launch_bounds__(kPostW * kPostH) k_postprocess(PostArgs a)
{
    __shared__ float4 s_tm[kTileH][kTileW];   // xyz = TM(colour), w = colour alpha
    const int W = a.W, H = a.H;
    const int tx = (int)threadIdx.x & 31, ty = (int)threadIdx.x >> 5;
    int rowBegin = a.row0, rowEnd = a.row1, blockRow = (int)blockIdx.y;
    if (a.stripeH) {
        const int blocksPerStripe = (a.stripeH + kPostH - 1) / kPostH;
        const int k = blockRow / blocksPerStripe;
        blockRow -= k * blocksPerStripe;
        rowBegin = (k * a.world + a.rank) * a.stripeH; rowEnd = min(rowBegin + a.stripeH, H);
    }
    const int y0Stripe = rowBegin;     // first row of the stripe / band this CTA works in
    const int x0 = (int)blockIdx.x * kPostW, y0 = rowBegin + blockRow * kPostH;
    const int x = x0 + tx, y = y0 + ty;
    const bool valid = x < W && y < rowEnd;

    if (!a.taaOn) {
        if (valid) {
            const size_t pix = (size_t)y * W + x;
            const uint2 t = __ldg(a.color + pix);
            a.out[pix] = t;
            if (a.peerAllRows || y == y0Stripe || y == rowEnd - 1)
                for (int p = 0; p < a.numPeers; ++p) if (a.peerOut[p]) a.peerOut[p][pix] = t;
            const uchar4 bb = tone_map(t, a.toneLut);
            a.backBuffer[pix] = bb;
            if (a.peerBackBuffer) a.peerBackBuffer[pix] = bb;
        }
        return;
    }

    for (int i = (int)threadIdx.x; i < kTileW * kTileH; i += kPostW * kPostH) {
        const int lx = i % kTileW, ly = i / kTileW;
        const V4 c = load_c(a.color, x0 + lx - 1, y0 + ly - 1, W, H);
        const V3 t = TM(V3{c.x, c.y, c.z});
        s_tm[ly][lx] = make_float4(t.x, t.y, t.z, c.w);
    }
    __syncthreads();
    if (!valid) return;
    const size_t pix = (size_t)y * W + x;

    const int offs[8][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {1, 1}, {-1, 1}};   // :46-50
    const float historyMax = 15.0f;                                                                   // :41-43
    const V2 texSize = {(float)W, (float)H};
    const float4 own = s_tm[ty + 1][tx + 1];
    // VelocityMax :133-161
    V2 vmax = {0.0f, 0.0f};
    if (a.velocityGiven) {      // (an all-zero field gives vmax = 0 whichever tap wins)
        vmax = load_v(a.velocity, x, y, W, H);
        float speedSq = fma1(vmax.x, vmax.x, vmax.y * vmax.y);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const V2 nb = load_v(a.velocity, x + offs[i + 4][0], y + offs[i + 4][1], W, H);
            const float sq = fma1(nb.x, nb.x, nb.y * nb.y);
            if (sq > speedSq) { vmax = nb; speedSq = sq; }
        }
    }
    // history.SampleLevel(g_smpLinear, uvBack, 0): bilinear, clamp. The texel coordinate is formed the way the texture unit
    // forms it — fixed point, 8 fractional bits (tex_coord_q8) — so that a fetch at a texel centre returns that texel and
    // the `historyBlur > 0` test below does not hang on the last ulp of u * W - 0.5; the blend itself is fp32.
    V4 history;
    if (!a.velocityGiven) history = unpack_half4(__ldg(a.history + pix));      // uvBack = the pixel's own centre: weights (1, 0, 0, 0)
    else {
        const V2 uv = {((float)x + 0.5f) / texSize.x, ((float)y + 0.5f) / texSize.y};   // :258, a division in the shipped DXIL too
        const V2 uvBack = {uv.x - vmax.x, uv.y - vmax.y};
        const int xq = tex_coord_q8(uvBack.x, W), yq = tex_coord_q8(uvBack.y, H);
        const float wx = (float)(xq & 255) * 0.00390625f, wy = (float)(yq & 255) * 0.00390625f;
        const int xa = xq >> 8, xb = min(xa + 1, W - 1);
        const int ya = yq >> 8, yb = min(ya + 1, H - 1);
        const uint2* rowA = a.history + (size_t)ya * W;
        const uint2* rowB = a.history + (size_t)yb * W;
        const V4 t00 = unpack_half4(__ldg(rowA + xa)), t10 = unpack_half4(__ldg(rowA + xb));
        const V4 t01 = unpack_half4(__ldg(rowB + xa)), t11 = unpack_half4(__ldg(rowB + xb));
        history = {lerp_q8(lerp_q8(t00.x, t10.x, wx), lerp_q8(t01.x, t11.x, wx), wy), lerp_q8(lerp_q8(t00.y, t10.y, wx), lerp_q8(t01.y, t11.y, wx), wy),
                   lerp_q8(lerp_q8(t00.z, t10.z, wx), lerp_q8(t01.z, t11.z, wx), wy), lerp_q8(lerp_q8(t00.w, t10.w, wx), lerp_q8(t01.w, t11.w, wx), wy)};
    }
    // :267-275
    float curHistoryBlur = fma1(fabsf(vmax.x), 4.0f * texSize.x, fabsf(vmax.y) * (4.0f * texSize.y));
    float historyBlur = 1.0f - history.w;
    historyBlur = fmaxf(historyBlur, curHistoryBlur);
    history.w = fma1(history.w, historyMax, 1.0f);
    // :278-287 (ALPHA_BOUND = 1.0)
    const V4 currentTM = {own.x, own.y, own.z, own.w};
    const float gamma = (historyBlur > 0.0f || own.w < 1.0f) ? 1.0f : 16.0f;
    // NeighborMinMax :166-236
    V4 cur = currentTM;
    V3 mu = {cur.x, cur.y, cur.z};
    cur.w = cur.w < 1.0f ? 0.0f : 1.0f;
    V3 m2 = mu * mu;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float wgt = i < 4 ? 0.5f : 0.25f;
        const float4 nb = s_tm[ty + 1 + offs[i][1]][tx + 1 + offs[i][0]];
        const float na = nb.w < 1.0f ? 0.0f : 1.0f;
        cur = {fma1(nb.x, wgt, cur.x), fma1(nb.y, wgt, cur.y), fma1(nb.z, wgt, cur.z), fma1(na, wgt, cur.w)};
        mu = {mu.x + nb.x, mu.y + nb.y, mu.z + nb.z};
        m2 = {fma1(nb.x, nb.x, m2.x), fma1(nb.y, nb.y, m2.y), fma1(nb.z, nb.z, m2.z)};
    }
    const float ninth = kNinth;
    cur = cur * 0.25f;
    mu = mu * ninth;
    const V3 m2n = m2 * ninth;
    const V3 sigma = {sqrtf(fabsf(fma1(-mu.x, mu.x, m2n.x))), sqrtf(fabsf(fma1(-mu.y, mu.y, m2n.y))), sqrtf(fabsf(fma1(-mu.z, mu.z, m2n.z)))};
    const V3 gsigma = sigma * gamma;
    V4 nmin, nmax;
    nmin.x = fminf(mu.x - gsigma.x, cur.x); nmin.y = fminf(mu.y - gsigma.y, cur.y); nmin.z = fminf(mu.z - gsigma.z, cur.z);
    nmax.x = fmaxf(mu.x + gsigma.x, cur.x); nmax.y = fmaxf(mu.y + gsigma.y, cur.y); nmax.z = fmaxf(mu.z + gsigma.z, cur.z);
    nmin.w = mu.x - sigma.x;   // GET_LUMA4 = .x in YCoCg
    nmax.w = mu.x + sigma.x;
    V4 filtered = cur;
    // :290-301
    curHistoryBlur = saturate(curHistoryBlur);
    historyBlur = saturate(historyBlur);
    V3 historyTM = TM(V3{history.x, history.y, history.z});
    historyTM = {fminf(fmaxf(historyTM.x, nmin.x), nmax.x), fminf(fmaxf(historyTM.y, nmin.y), nmax.y), fminf(fmaxf(historyTM.z, nmin.z), nmax.z)};
    const float contrast = nmax.w - nmin.w;
    // :304-311
    const float lumContrastFactor = 32.0f * 4.0f;
    float addAlias = fma1(historyBlur, 0.5f, 0.25f);
    addAlias = saturate(addAlias + rcp(fma1(contrast, lumContrastFactor, 1.0f)));
    filtered.x = lerpf(filtered.x, currentTM.x, addAlias); filtered.y = lerpf(filtered.y, currentTM.y, addAlias);
    filtered.z = lerpf(filtered.z, currentTM.z, addAlias);
    // :314-326
    const float lumHist = historyTM.x;
    const float distToClamp = fminf(fabsf(nmin.w - lumHist), fabsf(nmax.w - lumHist));
    const float historyAmt = fminf(fma1(historyBlur, 0.125f, rcp(history.w)), 1.0f);
    float blend = 0.25f * rcp(lerpf(8.0f, distToClamp + contrast, historyAmt));
    blend = fminf(blend, 0.25f);
    blend = filtered.w > 0.0f ? blend : 1.0f;
    // :328-330
    V3 result = ITM(V3{lerpf(historyTM.x, filtered.x, blend), lerpf(historyTM.y, filtered.y, blend), lerpf(historyTM.z, filtered.z, blend)});
    if (result.x != result.x || result.y != result.y || result.z != result.z) result = ITM(V3{filtered.x, filtered.y, filtered.z});
    history.w = fminf(history.w * (1.0f / historyMax), 1.0f - curHistoryBlur);
    const uint2 outTexel = pack_half4(V4{result.x, result.y, result.z, history.w});
    a.out[pix] = outTexel;
    if (a.peerAllRows || y == y0Stripe || y == rowEnd - 1)
        for (int p = 0; p < a.numPeers; ++p) if (a.peerOut[p]) a.peerOut[p][pix] = outTexel;
    const uchar4 bb = tone_map(outTexel, a.toneLut);
    a.backBuffer[pix] = bb;
    if (a.peerBackBuffer) a.peerBackBuffer[pix] = bb;
}

} // namespace

void build_tone_lut(Caster& c) { k_build_tone_lut<<<256, 256, 0, c.stream>>>(c.dToneLut); }

void launch_postprocess(Caster& c, bool taaOn)
{
    c.frameParity ^= 1u;                                   // ObjectRenderer.cpp:217
    PostArgs a;
    a.color = c.dColor;
    a.history = c.dHistory[c.frameParity ^ 1u];
    a.velocity = c.dVelocity;
    a.out = c.dHistory[c.frameParity];
    a.backBuffer = c.dBackBuffer;
    a.peerBackBuffer = c.dPeerBackBuffer;
    a.numPeers = (c.shardWorld > 1 && c.peersMapped) ? (int)c.shardWorld : 0;
    // The next frame's history fetch of a pixel lands at uv - velocity: with the velocity field all zero (none was ever given)
    // that is the pixel itself up to rounding, i.e. within one row, so only the border rows of every stripe need to reach the
    // peers; with a velocity field it can land anywhere, and every row goes out.
    a.peerAllRows = c.velocityGiven ? 1 : 0;
    a.velocityGiven = c.velocityGiven ? 1 : 0;
    a.invW = 1.0f / (float)c.d.width; a.invH = 1.0f / (float)c.d.height;
    for (int p = 0; p < kMaxPeers; ++p) a.peerOut[p] = p < a.numPeers ? c.peerHistory[p][c.frameParity] : nullptr;
    a.W = (int)c.d.width; a.H = (int)c.d.height;
    a.row0 = (int)c.row0; a.row1 = (int)c.row1;
    a.taaOn = taaOn ? 1 : 0;
    a.toneLut = c.dToneLut;
    const bool stripes = c.shardWorld > 1 && c.stripeH;
    a.stripeH = stripes ? (int)c.stripeH : 0; a.rank = (int)c.shardRank; a.world = (int)c.shardWorld;
    const uint32_t blockRows = stripes ? num_own_stripes(c.d.height, c.stripeH, c.shardRank, c.shardWorld) * ((c.stripeH + kPostH - 1) / kPostH)
                                       : (c.row1 - c.row0 + kPostH - 1) / kPostH;
    if (blockRows == 0) return;
    dim3 grid((c.d.width + kPostW - 1) / kPostW, blockRows);
    k_postprocess<<<grid, kPostW * kPostH, 0, c.stream>>>(a);
}

} // namespace mv
