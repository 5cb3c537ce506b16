"""The one-launch weight preparation (tbi_prepare_run / tbi_bn_fold_multi over the item tables of prep.py) must write exactly
what the per-layer packing entry points write (tbi_pack_conv_weights / tbi_pack_convt_weights / tbi_bn_fold), bit for bit, for
every layer of every configuration: plain, grouped, block-diagonal expanded (r4k4), expanded with a padded input width (r3k4),
transposed convs, the 16-channel-padded head gradient pack and the head's two GEMM packs."""
import pytest
import torch

from oracle import tbi_resnest_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("radix,kpaths,dtype", [(2, 1, "bf16"), (4, 4, "bf16"), (3, 4, "bf16"), (2, 1, "fp32"), (3, 4, "fp32")])
def test_prepare_tables_equal_per_layer_packing(cuda_device, radix, kpaths, dtype):
    from ultrasound_modeling_b200 import ops, _lib
    from ultrasound_modeling_b200.TBI_ResNest import ResNest
    net = ResNest(64, 64, 1, 3, 3, radix=radix, kpaths=kpaths, dtype=dtype, use_cuda_graph=False)
    net.load_state_dict(O.TBIResNestOracle(64, 64, 1, 3, 3, radix, kpaths, dtype=torch.float32).state_dict())      # perturbed BN statistics
    e = net.engine
    e.build(2)
    e.prepare(); e._join_prepare()
    torch.cuda.synchronize()
    L, td = e.L, e.tdtype
    checked = 0
    for Lr in e.convs.values():
        pk = e.packed[Lr.name]
        bn = tuple(t for t in (e.p(Lr.name + "/gamma"), e.p(Lr.name + "/beta"), e.s(Lr.name + "/mean"), e.s(Lr.name + "/var"))) if Lr.bn else None
        scale, fbias = ops.fold_bn(Lr.cout, e.p(Lr.name + "/b"), bn, e.device)
        assert torch.equal(fbias, pk["fbias"]), Lr.name
        if scale is not None:
            assert torch.equal(scale, pk["scale"]), Lr.name
        w = e.p(Lr.name + "/w")
        if Lr.kind == "conv":
            for mode, key in ((0, "wf"), (1, "wb")):
                want = ops.pack_conv(w, Lr.groups, mode, td, scale)
                assert torch.equal(want, pk[key]), (Lr.name, mode)
        else:
            assert torch.equal(ops.pack_convt(w, 0, td, scale), pk["wf"]), Lr.name
            cpad = e.dl_c if Lr is e.head else 0
            assert torch.equal(ops.pack_convt(w, 1, td, scale, cout_pad=cpad), pk["wb"]), Lr.name
            if Lr is e.head and e.head_gather:
                want = torch.empty_like(e.head_wg)
                _lib.check(L.tbi_pack_convt_weights(e.dt, 2, Lr.k, Lr.cin, Lr.cout, 64, w.data_ptr(), None, want.data_ptr(), e.stream()), "pack")
                assert torch.equal(want, e.head_wg)
            if Lr is e.head and e.head_fwd_gemm:
                want = torch.empty_like(e.head_wf)
                _lib.check(L.tbi_pack_convt_weights(e.dt, 3, Lr.k, Lr.cin, Lr.cout, e.head_yc, w.data_ptr(), None, want.data_ptr(), e.stream()), "pack")
                assert torch.equal(want, e.head_wf)
        checked += 1
    assert checked == len(e.convs) >= 20
    # a second preparation after the weights moved writes the new values everywhere (nothing is cached)
    e.params.mul_(1.5)
    e.prepare(); e._join_prepare()
    Lr = e.convs["upsample_2"]
    scale, _ = ops.fold_bn(Lr.cout, e.p(Lr.name + "/b"), (e.p(Lr.name + "/gamma"), e.p(Lr.name + "/beta"), e.s(Lr.name + "/mean"), e.s(Lr.name + "/var")), e.device)
    assert torch.equal(ops.pack_convt(e.p(Lr.name + "/w"), 1, td, scale), e.packed[Lr.name]["wb"])
