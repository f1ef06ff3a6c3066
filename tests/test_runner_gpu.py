"""The job runner (lssvc_b200/runner.py): the reference's frame loop + fan-out (test.py:121-325, 685-748) on coding lanes.

CPU: the issue schedule covers every frame of every unit exactly once, serial inside a unit, lanes in lock step.
GPU: a two-sequence job coded (a) by the plain frame loop through the public API, (b) by the runner on one eager lane,
(c) on two lanes replaying CUDA graphs, (d) on three eager lanes gives the SAME table: bits equal up to the order of the
double-precision atomic adds, squared errors bit-identical — concurrency changes nothing in a frame."""
import pytest
import torch

from lssvc_b200 import gop


def test_lane_schedule_covers_every_frame_once():
    from lssvc_b200.runner import lane_schedule
    units = gop.work_units(3, 10, 4)                      # 9 units, ragged last GOPs
    for lanes in (1, 2, 4, 16):
        order = lane_schedule(units, lanes)
        seen = [(u.seq, f) for _, u, f, _ in order]
        assert len(seen) == len(set(seen)) == 30
        for k in range(lanes):                              # a lane's frames: unit after unit, frames in coding order
            mine = [(u, f, i) for lane, u, f, i in order if lane == k]
            want = [(u, f, i) for u in units[k::lanes] for f, i in gop.frames_of(u)]
            assert mine == want
        for u in units:
            assert [i for _, uu, f, i in order if uu == u] == [True] + [False] * (u.n_frames - 1)
    assert lane_schedule([], 3) == []


@pytest.mark.gpu
def test_runner_lanes_reproduce_the_frame_loop(cuda_device):
    from lssvc_b200 import IntraSS, LSSVC_extend, frontend, synth
    from lssvc_b200.runner import GopRunner
    dev = cuda_device
    H = W = 256
    n_seq, n_frames, gop_size = 2, 9, 4                    # units of 4, 4 and 1 frames per sequence
    net_i = IntraSS(seed=0).to(dev)
    net_p = LSSVC_extend(seed=1).to(dev)
    for n in (net_i, net_p):
        n.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
    seqs = [[(b.to(dev), e.to(dev)) for b, e in synth.make_sequence(H, W, n_frames, seed=10 + s)] for s in range(n_seq)]
    source = lambda seq, f: seqs[seq][f]
    units = gop.work_units(n_seq, n_frames, gop_size)

    # (a) the frame loop of test.py through the public API
    want = []
    for s in range(n_seq):
        dpb = None
        for f in range(n_frames):
            x_bl, x_el = seqs[s][f]
            if f % gop_size == 0:
                r = net_i.encode_decode(x_bl, x_el, None, None, H // 2, W // 2, H, W)
                dpb = {"ref_frame_bl": r["x_hat_bl"], "ref_frame_el": r["x_hat_el"], "ref_feature_bl": None, "ref_feature_el": r["feature_el"]}
            else:
                r = net_p.encode_decode(x_bl, x_el, dpb, None, None, W, H, W // 2, H // 2)
                dpb = r["dpb"]
            rb, re = dpb["ref_frame_bl"].clamp_(0, 1), dpb["ref_frame_el"].clamp_(0, 1)
            sse = [float(((a - b).double() ** 2).sum()) for a, b in ((x_bl, rb), (x_el, re))]
            want.append((s, f, float(f % gop_size == 0), r["bit_bl"], r["bit_el"], sse[0], sse[1]))
            assert abs(frontend.psnr(x_el, re) - 10 * torch.log10(torch.tensor(x_el.numel() / sse[1])).item()) < 1e-6
    want = torch.tensor(want, dtype=torch.float64)

    first = None
    for lanes, graphs in ((1, False), (2, True), (3, False), (2, None)):
        runner = GopRunner(net_i, net_p, lanes=lanes, graphs=graphs)
        rows = runner.code_units(units, source)
        runner.synchronize()
        got = gop.gather_stats(rows)
        assert got.shape == want.shape == (n_seq * n_frames, len(gop.STAT_COLUMNS))
        assert torch.equal(got[:, :3], want[:, :3])
        rel = ((got[:, 3:5] - want[:, 3:5]).abs() / want[:, 3:5]).max().item()
        d_sse = ((got[:, 5:] - want[:, 5:]).abs() / want[:, 5:]).max().item()
        # against the plain loop: the bits are the same kernels (atomics order only); the reference SSE above is torch's
        # double-precision sum of fp32 squares, the kernel squares in double: 1e-8.  Between runner configurations: identical.
        first = got if first is None else first
        same = ((got[:, 3:] - first[:, 3:]).abs() / first[:, 3:]).max().item()
        print(f"lanes={lanes} graphs={graphs}: bits rel {rel:.1e}, sse rel {d_sse:.1e}; against the one-lane runner {same:.1e}")
        assert rel < 1e-9 and d_sse < 1e-8 and same < 1e-12
        if graphs:
            assert any(isinstance(g, dict) for g in net_p._graphs.values()), "no frame graph was captured"
    s = gop.summarize(want, H * W, H * W // 4)
    assert s["frames"] == n_seq * n_frames and s["psnr_el"] > 10 and s["bpp_el"] > 0
