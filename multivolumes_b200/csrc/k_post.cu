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
namespace MV_VARIANT {

namespace {

MV_D V3 rgb_to_ycocg(V3 rgb)   // :78-85
{
    const float y = (rgb.x * 1.0f + rgb.y * 2.0f) + rgb.z * 1.0f;
    const float co = (rgb.x * 2.0f + rgb.y * 0.0f) + rgb.z * -2.0f;
    const float cg = (rgb.x * -1.0f + rgb.y * 2.0f) + rgb.z * -1.0f;
    return {y, co, cg};
}
MV_D V3 ycocg_to_rgb(V3 v)     // :90-101
{
    const float y = v.x * 0.25f, co = v.y * 0.25f, cg = v.z * 0.25f;
    return {y + co - cg, y + cg, y - co - cg};
}
MV_D V3 TM(V3 hdr) { const V3 c = rgb_to_ycocg(hdr); const float d = 4.0f + c.x; return {c.x / d, c.y / d, c.z / d}; }   // :106-114
MV_D V3 ITM(V3 c) { const float s = 4.0f / (1.0f - c.x); return ycocg_to_rgb(V3{c.x * s, c.y * s, c.z * s}); }         // :119-128

MV_D uchar4 tone_map(uint2 taaTexel)   // PSToneMap.hlsl:19-28 + RGBA8_UNORM render-target write
{
    const V4 src = unpack_half4(taaTexel);
    const float in[3] = {src.x, src.y, src.z};
    unsigned char out[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = in[k];
        v *= 1.05f / (v + 0.7f);
        v = pow125(fabsf(v));
        float sat = saturate(v);
        if (!(v == v)) sat = 0.0f;
        out[k] = (unsigned char)floorf(sat * 255.0f + 0.5f);
    }
    return make_uchar4(out[0], out[1], out[2], 255);
}

struct PostArgs {
    const uint2* color;       // current frame, RGBA16F
    const uint2* history;     // previous TAA output
    const uint32_t* velocity; // RG16F
    uint2* out;               // TAA output (next frame's history)
    uchar4* backBuffer;       // RGBA8
    uchar4* peerBackBuffer;   // rank 0's back buffer when this rank resolves a band of a multi-GPU frame
    int W, H, row0, row1;
    int stripeH, rank, world;   // stripeH > 0: interleaved stripes instead of the band
    int taaOn;
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
    const int x0 = (int)blockIdx.x * kPostW, y0 = rowBegin + blockRow * kPostH;
    const int x = x0 + tx, y = y0 + ty;
    const bool valid = x < W && y < rowEnd;

    if (!a.taaOn) {
        if (valid) {
            const size_t pix = (size_t)y * W + x;
            const uint2 t = __ldg(a.color + pix);
            a.out[pix] = t;
            const uchar4 bb = tone_map(t);
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
    const V2 uv = {((float)x + 0.5f) / texSize.x, ((float)y + 0.5f) / texSize.y};
    const float4 own = s_tm[ty + 1][tx + 1];
    // VelocityMax :133-161
    V2 vmax = load_v(a.velocity, x, y, W, H);
    float speedSq = dot(vmax, vmax);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const V2 nb = load_v(a.velocity, x + offs[i + 4][0], y + offs[i + 4][1], W, H);
        const float sq = dot(nb, nb);
        if (sq > speedSq) { vmax = nb; speedSq = sq; }
    }
    const V2 uvBack = {uv.x - vmax.x, uv.y - vmax.y};
    // history.SampleLevel(g_smpLinear, uvBack, 0): bilinear, clamp, fp32 weights
    V4 history;
    {
        const float fx = uvBack.x * texSize.x - 0.5f, fy = uvBack.y * texSize.y - 0.5f;
        const float flx = floorf(fx), fly = floorf(fy);
        const float wx = fx - flx, wy = fy - fly;
        const int ix = (int)flx, iy = (int)fly;
        const int xa = min(max(ix, 0), W - 1), xb = min(max(ix + 1, 0), W - 1);
        const int ya = min(max(iy, 0), H - 1), yb = min(max(iy + 1, 0), H - 1);
        const V4 t00 = load_c(a.history, xa, ya, W, H), t10 = load_c(a.history, xb, ya, W, H);
        const V4 t01 = load_c(a.history, xa, yb, W, H), t11 = load_c(a.history, xb, yb, W, H);
        history = {lerp(lerp(t00.x, t10.x, wx), lerp(t01.x, t11.x, wx), wy), lerp(lerp(t00.y, t10.y, wx), lerp(t01.y, t11.y, wx), wy),
                   lerp(lerp(t00.z, t10.z, wx), lerp(t01.z, t11.z, wx), wy), lerp(lerp(t00.w, t10.w, wx), lerp(t01.w, t11.w, wx), wy)};
    }
    // :267-275
    const V2 historyBlurAmp = {4.0f * texSize.x, 4.0f * texSize.y};
    const V2 historyBlurs = {fabsf(vmax.x) * historyBlurAmp.x, fabsf(vmax.y) * historyBlurAmp.y};
    float curHistoryBlur = historyBlurs.x + historyBlurs.y;
    float historyBlur = 1.0f - history.w;
    historyBlur = fmaxf(historyBlur, curHistoryBlur);
    history.w = history.w * historyMax + 1.0f;
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
        const V3 ntm = {nb.x, nb.y, nb.z};
        const V4 neighbor = {ntm.x, ntm.y, ntm.z, nb.w < 1.0f ? 0.0f : 1.0f};
        cur = cur + neighbor * wgt;
        mu = mu + ntm;
        m2 = m2 + ntm * ntm;
    }
    cur = {cur.x / 4.0f, cur.y / 4.0f, cur.z / 4.0f, cur.w / 4.0f};
    mu = mu / 9.0f;
    const V3 m2n = m2 / 9.0f;
    const V3 sigma = {sqrtf(fabsf(m2n.x - mu.x * mu.x)), sqrtf(fabsf(m2n.y - mu.y * mu.y)), sqrtf(fabsf(m2n.z - mu.z * mu.z))};
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
    float addAlias = historyBlur * 0.5f + 0.25f;
    addAlias = saturate(addAlias + 1.0f / (1.0f + contrast * lumContrastFactor));
    filtered.x = lerp(filtered.x, currentTM.x, addAlias); filtered.y = lerp(filtered.y, currentTM.y, addAlias);
    filtered.z = lerp(filtered.z, currentTM.z, addAlias);
    // :314-326
    const float lumHist = historyTM.x;
    const float distToClamp = fminf(fabsf(nmin.w - lumHist), fabsf(nmax.w - lumHist));
    const float historyAmt = fminf(1.0f / history.w + historyBlur / 8.0f, 1.0f);
    float blend = 0.25f / lerp(8.0f, distToClamp + contrast, historyAmt);
    blend = fminf(blend, 0.25f);
    blend = filtered.w > 0.0f ? blend : 1.0f;
    // :328-330
    V3 result = ITM(V3{lerp(historyTM.x, filtered.x, blend), lerp(historyTM.y, filtered.y, blend), lerp(historyTM.z, filtered.z, blend)});
    if (result.x != result.x || result.y != result.y || result.z != result.z) result = ITM(V3{filtered.x, filtered.y, filtered.z});
    history.w = fminf(history.w / historyMax, 1.0f - curHistoryBlur);
    const uint2 outTexel = pack_half4(V4{result.x, result.y, result.z, history.w});
    a.out[pix] = outTexel;
    const uchar4 bb = tone_map(outTexel);
    a.backBuffer[pix] = bb;
    if (a.peerBackBuffer) a.peerBackBuffer[pix] = bb;
}

} // namespace

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
    a.W = (int)c.d.width; a.H = (int)c.d.height;
    a.row0 = (int)c.row0; a.row1 = (int)c.row1;
    a.taaOn = taaOn ? 1 : 0;
    const bool stripes = c.shardWorld > 1 && c.stripeH;
    a.stripeH = stripes ? (int)c.stripeH : 0; a.rank = (int)c.shardRank; a.world = (int)c.shardWorld;
    const uint32_t blockRows = stripes ? num_own_stripes(c.d.height, c.stripeH, c.shardRank, c.shardWorld) * ((c.stripeH + kPostH - 1) / kPostH)
                                       : (c.row1 - c.row0 + kPostH - 1) / kPostH;
    if (blockRows == 0) return;
    dim3 grid((c.d.width + kPostW - 1) / kPostW, blockRows);
    k_postprocess<<<grid, kPostW * kPostH, 0, c.stream>>>(a);
}

} // namespace MV_VARIANT
} // namespace mv
