"""GPU: hwg_linear_map (hwg_map.cu) against the CPU interpreter of the same job tables, and the generator's packed
operands against the torch re-layouts of the reference definitions."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_linear_map_kernel_matches_cpu_interpreter():
    from handwriting_line_generation_b200 import weightmap as wm
    from tests import ref_map
    torch.manual_seed(0)
    maps = [wm.map_conv3x3(32, 16), wm.map_initial(22, 32), wm.map_vert_up(16, 32), wm.map_fused_up(32, 16, 0.0589)]
    shapes = [(32, 16, 3, 3), (22, 32, 4, 3), (16, 32, 3, 3), (32, 16, 3, 3)]
    tc, tg = wm.JobTable(), wm.JobTable()
    cpu_out, gpu_out = [], []
    ws_c = torch.randn(40000)
    gf_c = torch.full((30000,), 3.0)
    ws_g, gf_g = ws_c.cuda(), gf_c.cuda()
    wo, go = 0, 0
    for m, sh in zip(maps, shapes):
        w = torch.randn(sh)
        wg = w.cuda()
        for t, wt, outs, dev in ((tc, w, cpu_out, "cpu"), (tg, wg, gpu_out, "cuda")):
            pf = torch.full((m.Tf, m.Co, m.Cip), 9.0, dtype=torch.bfloat16, device=dev)
            pd = torch.full(m.dgrad_shape(), 9.0, dtype=torch.bfloat16, device=dev)
            m.add_pack_fwd(t, wt, pf)
            m.add_pack_dgrad(t, wt, pd)
            m.add_unpack_wgrad(t, 4 * wo, 4 * go, accumulate=(sh[0] == 22))   # relative to the ws / gflat bases
            outs += [pf, pd]
        wo += m.Tf * m.Co * m.Cip
        go += w.numel()
    for t in (tc, tg):   # a batch-sum job and a scaled vector job
        t.add(4 * 100, 4 * (go + 8), R=1, C=8, s_r=0, s_c=2, d_r=0, d_c=1, M=None, nin=5, in_stride=16, scale=0.5)
        t.add(4 * 7, 4 * go, R=1, C=8, s_r=0, s_c=1, d_r=0, d_c=1, M=[[1.0, 1.0]], out_off=[0, 20], scale=0.25)
    ref_map.run_jobs_cpu(tc, src_base=ws_c, dst_base=gf_c)
    tg.finalize(torch.device("cuda"))
    tg.run(src_base=ws_g, dst_base=gf_g)
    torch.cuda.synchronize()
    for a, b in zip(cpu_out, gpu_out):
        assert torch.equal(a.float(), b.float().cpu()) or (a.float() - b.float().cpu()).abs().max() <= 1e-2 * a.float().abs().max()
    assert torch.allclose(gf_c, gf_g.cpu(), atol=1e-5, rtol=1e-5)


def test_generator_packed_operands_match_torch_relayouts():
    import handwriting_line_generation_b200 as pkg
    from tests import ref_pack
    torch.manual_seed(3)
    g = pkg.SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True).cuda()
    c = g._packed()
    torch.cuda.synchronize()

    def close(a, b):
        a, b = a.float(), b.float()
        assert a.shape == b.shape, (a.shape, b.shape)
        assert (a - b).abs().max() <= 8e-3 * b.abs().max() + 1e-12

    b0, b1, b3 = c["blocks"][0], c["blocks"][1], c["blocks"][3]
    w = g.conv[0].conv1.weight.detach()
    close(b0["w1f"], ref_pack.initial_fwd(w, c["cin_pad"]))
    close(b0["d1"], ref_pack.initial_dgrad(w))
    close(b0["b1f"], g.conv[0].conv1.bias.detach().repeat(4))
    close(b0["nw1f"], g.conv[0].noise1.effective_weight().detach().repeat(4))
    w = g.conv[1].conv1[1].weight.detach()
    fw = ref_pack.vert_up_fwd(w)
    for par in (0, 1):
        close(b1["w1"][par][1], fw[par])
    close(b1["d1"], ref_pack.vert_up_dgrad(w))
    mod = g.conv[3].conv1[0]
    close(b3["w1f"], ref_pack.fused_up_fwd(mod.weight.detach(), mod.multiplier)[0])
    close(b3["d1"], ref_pack.fused_up_dgrad(mod.weight.detach(), mod.multiplier))
    for bi, e in enumerate(c["blocks"]):
        w2 = g.conv[bi].conv2.weight.detach()
        close(e["w2"], ref_pack.conv3x3_fwd(w2))
        close(e["d2"], ref_pack.conv3x3_dgrad(w2))
        close(e["nw2"], g.conv[bi].noise2.effective_weight().detach())
    close(c["w_out"], g.out[0].effective_weight().detach().reshape(-1))
    ws = torch.cat([ad.style.weight for blk in g.conv for ad in (blk.adain1, blk.adain2)], 0).detach()
    bs = torch.cat([ad.style.bias for blk in g.conv for ad in (blk.adain1, blk.adain2)], 0).detach()
    close(c["gb_w"], ws)
    close(c["gb_b"], bs)
    # in-place re-pack after a parameter update
    with torch.no_grad():
        g.conv[2].conv2.weight.mul_(0.5)
    c2 = g._packed()
    close(c2["blocks"][2]["w2"], ref_pack.conv3x3_fwd(g.conv[2].conv2.weight.detach()))
