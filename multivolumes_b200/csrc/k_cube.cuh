// k_cube.cuh — cube-map addressing (D3D face convention, seamless edge taps) and the row ownership of a sharded frame, shared
// by the OIT resolve (k_oit.cu) and the environment pass (k_env.cu).
#pragma once
#include "mv_internal.h"

namespace mv {

// D3D cube-map convention: (u, v) in [0, 1] of point p on face `face` of the unit cube
MV_D void cube_face_uv(V3 p, int face, float& u, float& v)
{
    switch (face) {
    case 0: u = -p.z * 0.5f + 0.5f; v = -p.y * 0.5f + 0.5f; break;
    case 1: u = p.z * 0.5f + 0.5f; v = -p.y * 0.5f + 0.5f; break;
    case 2: u = p.x * 0.5f + 0.5f; v = p.z * 0.5f + 0.5f; break;
    case 3: u = p.x * 0.5f + 0.5f; v = -p.z * 0.5f + 0.5f; break;
    case 4: u = p.x * 0.5f + 0.5f; v = -p.y * 0.5f + 0.5f; break;
    default: u = -p.x * 0.5f + 0.5f; v = -p.y * 0.5f + 0.5f; break;
    }
}

// Seamless cube addressing for the Gather* taps of CubeCast: texel (i, j) of `face`, where an index
// may be -1 or S, resolved to the edge-adjacent texel of the neighbouring face (a corner is pinned to
// the edge texel). Integer arithmetic on odd coordinates in units of 1 / S.
MV_D void cube_resolve_texel(int S, int face, int i, int j, int& oface, int& oi, int& oj)
{
    const bool iOut = i < 0 || i >= S;
    if (iOut && (j < 0 || j >= S)) j = j < 0 ? 0 : S - 1;
    const bool jOut = j < 0 || j >= S;
    if (!iOut && !jOut) { oface = face; oi = i; oj = j; return; }
    const int a = 2 * i + 1 - S, b = 2 * j + 1 - S;
    int P[3];
    switch (face) {
    case 0: P[0] = S; P[1] = -b; P[2] = -a; break;
    case 1: P[0] = -S; P[1] = -b; P[2] = a; break;
    case 2: P[0] = a; P[1] = S; P[2] = b; break;
    case 3: P[0] = a; P[1] = -S; P[2] = -b; break;
    case 4: P[0] = a; P[1] = -b; P[2] = S; break;
    default: P[0] = -a; P[1] = -b; P[2] = -S; break;
    }
    const int major = face >> 1;
    int over = -1;
#pragma unroll
    for (int k = 0; k < 3; ++k) if (k != major && (P[k] > S || P[k] < -S)) over = k;
    // exactly one in-plane coordinate is outside by one half-texel step: fold it onto the neighbour
    const int pm = major == 0 ? P[0] : (major == 1 ? P[1] : P[2]);
    const int po = over == 0 ? P[0] : (over == 1 ? P[1] : P[2]);
    const int e = (po > 0 ? po : -po) - S;
    const int newMajor = (pm > 0 ? 1 : -1) * (S - e);
    const int newOver = po > 0 ? S : -S;
#pragma unroll
    for (int k = 0; k < 3; ++k) { if (k == major) P[k] = newMajor; if (k == over) P[k] = newOver; }
    oface = over * 2 + (newOver > 0 ? 0 : 1);
    int ua, vb;
    switch (oface) {
    case 0: ua = -P[2]; vb = -P[1]; break;
    case 1: ua = P[2]; vb = -P[1]; break;
    case 2: ua = P[0]; vb = P[2]; break;
    case 3: ua = P[0]; vb = -P[2]; break;
    case 4: ua = P[0]; vb = -P[1]; break;
    default: ua = -P[0]; vb = -P[1]; break;
    }
    oi = (ua + S - 1) / 2; oj = (vb + S - 1) / 2;
}

// Does this rank resolve row py (its band or stripes, plus the one-row halo the TAA's 3x3 neighbourhood reads)?
MV_D bool row_is_resolved_here(const DeviceScene& s, const FrameCB& cb, int py)
{
    if (py < 0 || py >= (int)cb.height) return false;
    if (!s.stripeH) return py >= (int)s.row0 - 1 && py < (int)s.row1 + 1;
    const int h = (int)s.stripeH, world = (int)s.shardWorld, rank = (int)s.shardRank;
    const int g = py / h;
    if (g % world == rank) return true;
    if (py == g * h && g >= 1 && (g - 1) % world == rank) return true;          // halo below the previous stripe
    if (py == (g + 1) * h - 1 && (g + 1) % world == rank && (g + 1) * h < (int)cb.height) return true;   // halo above the next stripe
    return false;
}

} // namespace mv
