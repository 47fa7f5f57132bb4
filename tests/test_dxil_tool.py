"""The tooling that runs the reference's compiled shaders (oracle/dxil) on hand-written IR: no reference file is needed here.
What is checked is what the golden vectors rely on: the wave scheduler's reconvergence, WaveReadLaneAt from a lane that is
switched off, phi semantics, integer / float conversions, min-precision promotion, resource access and the bitstream walker."""
import numpy as np
import pytest

from oracle.dxil import interp as I
from oracle.dxil.interp import CBuffer, Resources, Shader, StructuredBuffer, Texture, TypedBuffer, run_compute

HEAD = """
%dx.types.Handle = type { ptr }
%dx.types.ResRet.i32 = type { i32, i32, i32, i32, i32 }
%dx.types.ResRet.f32 = type { float, float, float, float, i32 }
"""


def _run(body, n=32, uav=None, srv=None, cbv=None, groups=(1, 1, 1), tpg=None):
    sh = Shader(HEAD + "define void @main() {\n" + body + "\n}\n")
    run_compute(sh, Resources(srv=srv, uav=uav, cbv=cbv), groups, threads_per_group=tpg or (n, 1, 1))
    return sh


def test_wave_sum_after_a_loop_with_per_lane_trip_counts_takes_every_lane():
    """lanes leave the loop after lane_index iterations; the WaveActiveSum behind it must wait for all of them"""
    out = StructuredBuffer(np.zeros(32, np.uint32), 4)
    _run("""
  %h = call %dx.types.Handle @dx.op.createHandle(i32 57, i8 1, i32 0, i32 0, i1 false)
  %lane = call i32 @dx.op.waveGetLaneIndex(i32 111)
  br label %loop
loop:
  %i = phi i32 [ 0, %0 ], [ %i1, %loop ]
  %acc = phi i32 [ 0, %0 ], [ %acc1, %loop ]
  %acc1 = add i32 %acc, 2
  %i1 = add i32 %i, 1
  %more = icmp ult i32 %i1, %lane
  br i1 %more, label %loop, label %done
done:
  %s = call i32 @dx.op.waveActiveOp.i32(i32 119, i32 %acc1, i8 0, i8 1)
  call void @dx.op.rawBufferStore.i32(i32 140, %dx.types.Handle %h, i32 %lane, i32 0, i32 %s, i32 undef, i32 undef, i32 undef, i8 1, i32 4)
  ret void""", uav={0: out})
    want = sum(2 * max(l, 1) for l in range(32))
    assert np.all(out.words == want)


def test_wave_op_reached_in_different_loop_iterations_groups_by_iteration():
    """VolumeCull.hlsli:250-257: `for (i < groups) if (i == lane / 8) return WaveActiveMax(x)` — one reduction per 8-lane group"""
    out = StructuredBuffer(np.zeros(32, np.uint32), 4)
    _run("""
  %h = call %dx.types.Handle @dx.op.createHandle(i32 57, i8 1, i32 0, i32 0, i1 false)
  %lane = call i32 @dx.op.waveGetLaneIndex(i32 111)
  %grp = lshr i32 %lane, 3
  br label %loop
loop:
  %i = phi i32 [ 0, %0 ], [ %i1, %next ]
  %mine = icmp eq i32 %i, %grp
  br i1 %mine, label %reduce, label %next
reduce:
  %m = call i32 @dx.op.waveActiveOp.i32(i32 119, i32 %lane, i8 3, i8 1)
  br label %done
next:
  %i1 = add i32 %i, 1
  %more = icmp ult i32 %i1, 4
  br i1 %more, label %loop, label %none
none:
  br label %done
done:
  %r = phi i32 [ %m, %reduce ], [ 999, %none ]
  call void @dx.op.rawBufferStore.i32(i32 140, %dx.types.Handle %h, i32 %lane, i32 0, i32 %r, i32 undef, i32 undef, i32 undef, i8 1, i32 4)
  ret void""", uav={0: out})
    assert np.array_equal(out.words, np.repeat([7, 15, 23, 31], 8))


def test_read_lane_at_serves_a_lane_that_is_switched_off():
    """VolumeCull.hlsli:153-154 reads the cube's vertices 6 and 7 from lanes the `wTidx < 6` branch has switched off"""
    out = StructuredBuffer(np.zeros(8, np.uint32), 4)
    _run("""
  %h = call %dx.types.Handle @dx.op.createHandle(i32 57, i8 1, i32 0, i32 0, i1 false)
  %lane = call i32 @dx.op.waveGetLaneIndex(i32 111)
  %v = mul i32 %lane, 10
  %on = icmp ult i32 %lane, 6
  br i1 %on, label %read, label %end
read:
  %src = sub i32 7, %lane
  %got = call i32 @dx.op.waveReadLaneAt.i32(i32 117, i32 %v, i32 %src)
  call void @dx.op.rawBufferStore.i32(i32 140, %dx.types.Handle %h, i32 %lane, i32 0, i32 %got, i32 undef, i32 undef, i32 undef, i8 1, i32 4)
  br label %end
end:
  ret void""", n=8, uav={0: out})
    assert list(out.words) == [70, 60, 50, 40, 30, 20, 0, 0]


def test_phis_of_a_block_read_their_inputs_simultaneously_and_ballot_prefix_count():
    out = StructuredBuffer(np.zeros((4, 4), np.uint32), 16)
    _run("""
  %h = call %dx.types.Handle @dx.op.createHandle(i32 57, i8 1, i32 0, i32 0, i1 false)
  %lane = call i32 @dx.op.waveGetLaneIndex(i32 111)
  br label %loop
loop:
  %a = phi i32 [ 1, %0 ], [ %b, %loop ]
  %b = phi i32 [ 2, %0 ], [ %a, %loop ]
  %i = phi i32 [ 0, %0 ], [ %i1, %loop ]
  %i1 = add i32 %i, 1
  %more = icmp ult i32 %i1, 3
  br i1 %more, label %loop, label %done
done:
  %odd = and i32 %lane, 1
  %p = icmp ne i32 %odd, 0
  %bal = call %dx.types.ResRet.i32 @dx.op.waveActiveBallot(i32 116, i1 %p)
  %mask = extractvalue %dx.types.ResRet.i32 %bal, 0
  %pre = call i32 @dx.op.wavePrefixOp.i32(i32 121, i32 %lane, i8 0, i8 1)
  call void @dx.op.rawBufferStore.i32(i32 140, %dx.types.Handle %h, i32 %lane, i32 0, i32 %a, i32 %b, i32 %mask, i32 %pre, i8 15, i32 4)
  ret void""", n=4, uav={0: out})
    w = out.words.reshape(4, 4)
    assert np.all(w[:, 0] == 1) and np.all(w[:, 1] == 2)          # swapped twice
    assert np.all(w[:, 2] == 0b1010) and list(w[:, 3]) == [0, 0, 1, 3]


def test_arithmetic_conversions_and_min_precision():
    out = StructuredBuffer(np.zeros(8, np.float32), 4)
    body = """
  %h = call %dx.types.Handle @dx.op.createHandle(i32 57, i8 1, i32 0, i32 0, i1 false)
  %neg = fptoui float -3.500000e+00 to i32
  %big = fptoui float 0x4202A05F20000000 to i32
  %sx = ashr i32 -16, 2
  %h1 = fptrunc float 0x3FB99999A0000000 to half
  %h2 = fmul fast half %h1, 0xH4900
  %f2 = fpext half %h2 to float
  %mad = call float @dx.op.tertiary.f32(i32 46, float 0x3FB99999A0000000, float 3.000000e+00, float -0x3FD3333340000000)
  %rs = call float @dx.op.unary.f32(i32 25, float 4.000000e+00)
  %d3 = call float @dx.op.dot3.f32(i32 55, float 1.0, float 2.0, float 3.0, float 4.0, float 5.0, float 6.0)
  %a = uitofp i32 %neg to float
  %b = uitofp i32 %big to float
  %c = sitofp i32 %sx to float
  call void @dx.op.rawBufferStore.f32(i32 140, %dx.types.Handle %h, i32 0, i32 0, float %a, float %b, float %c, float %f2, i8 15, i32 4)
  call void @dx.op.rawBufferStore.f32(i32 140, %dx.types.Handle %h, i32 4, i32 0, float %mad, float %rs, float %d3, float undef, i8 7, i32 4)
  ret void"""
    body = body.replace("-0x3FD3333340000000", "0xBFD3333340000000")
    res = {}
    for promote in (True, False):
        I.PROMOTE_HALF = promote
        out.words[:] = 0
        _run(body, n=1, uav={0: out})
        res[promote] = out.words.view(np.float32).copy()
    I.PROMOTE_HALF = True
    f = res[True]
    assert f[0] == 0.0 and f[1] == np.float32(4294967295.0) and f[2] == -4.0          # fptoui saturates, ashr keeps the sign
    assert f[3] == np.float32(0.1) * np.float32(10.0)                                   # min16 promoted: binary32 arithmetic, binary16 literal
    assert res[False][3] == np.float32(np.float16(np.float16(np.float32(0.1)) * np.float16(10.0)))
    assert f[4] == np.float32(np.float32(np.float32(0.1) * np.float32(3.0)) + np.float32(-0.3))   # FMad unfused
    assert f[5] == 0.5 and f[6] == 32.0


def test_resources_typed_buffer_texture_counter_and_cbuffer():
    tb = TypedBuffer(np.arange(12, dtype=np.uint32).reshape(3, 4))
    tex = Texture(np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4))               # [y][x][c]
    app = StructuredBuffer(np.zeros(8, np.uint32), 4)
    cb = CBuffer(np.array([0, 0, 0, 0, 5, 6, 7, 8], np.uint32).tobytes())
    _run("""
  %u = call %dx.types.Handle @dx.op.createHandle(i32 57, i8 1, i32 0, i32 0, i1 false)
  %t = call %dx.types.Handle @dx.op.createHandle(i32 57, i8 0, i32 0, i32 0, i1 false)
  %x = call %dx.types.Handle @dx.op.createHandle(i32 57, i8 0, i32 1, i32 1, i1 false)
  %c = call %dx.types.Handle @dx.op.createHandle(i32 57, i8 2, i32 0, i32 0, i1 false)
  %lane = call i32 @dx.op.waveGetLaneIndex(i32 111)
  %row = call %dx.types.ResRet.i32 @dx.op.bufferLoad.i32(i32 68, %dx.types.Handle %t, i32 %lane, i32 undef)
  %w = extractvalue %dx.types.ResRet.i32 %row, 3
  %tx = call %dx.types.ResRet.f32 @dx.op.textureLoad.f32(i32 66, %dx.types.Handle %x, i32 0, i32 %lane, i32 1, i32 undef, i32 undef, i32 undef, i32 undef)
  %g = extractvalue %dx.types.ResRet.f32 %tx, 1
  %gi = fptoui float %g to i32
  %k = call %dx.types.ResRet.i32 @dx.op.cbufferLoadLegacy.i32(i32 59, %dx.types.Handle %c, i32 1)
  %k2 = extractvalue %dx.types.ResRet.i32 %k, 2
  %sum = add i32 %w, %gi
  %sum2 = add i32 %sum, %k2
  %slot = call i32 @dx.op.bufferUpdateCounter(i32 70, %dx.types.Handle %u, i8 1)
  call void @dx.op.bufferStore.i32(i32 69, %dx.types.Handle %u, i32 %slot, i32 0, i32 %sum2, i32 undef, i32 undef, i32 undef, i8 1)
  ret void""", n=3, uav={0: app}, srv={0: tb, 1: tex}, cbv={0: cb})
    assert app.counter == 3
    # lane l: tb[l][3] = 4 l + 3; tex[y = 1][x = l][c = 1] = 12 + 4 l + 1; cb row 1 .z = 7
    assert sorted(app.words[:3]) == sorted(4 * l + 3 + 13 + 4 * l + 7 for l in range(3))


def test_bitstream_walker_finds_the_module_records():
    llvm = pytest.importorskip("llvmlite.binding")
    from oracle.dxil.bitstream import walk
    m = llvm.parse_assembly('target datalayout = "e-m:e-p:32:32-i64:64-n8:16:32"\ntarget triple = "dxil-ms-dx"\ndefine void @main() {\n  ret void\n}\n')
    recs = walk(m.as_bitcode())
    strings = {code: "".join(map(chr, ops)) for path, code, ops, _, _ in recs if path == (8,) and code in (2, 3)}
    assert strings[2] == "dxil-ms-dx" and strings[3] == "e-m:e-p:32:32-i64:64-n8:16:32"
    # the function block and its single `ret` (FUNC_CODE_INST_RET = 10) are reached
    assert any(path == (8, 12) and code == 10 for path, code, _, _, _ in recs)
