"""GPU (one device): the expert-parallel dispatch / layout / return kernels of csrc/ep_p2p.cu with W ranks EMULATED on
one GPU — every "rank" owns its own buffers on the same device, the peer-pointer arrays point at them, and the
cross-rank barriers become stream order (the kernels never wait on one another).  Checked against the numpy layout
oracle (bit-exact maps) and, end to end, against the unsharded MOELayer on the concatenated batch (the
reference's algorithm; moe_layer.py:122-173 has no expert parallelism, only the placeholder moe_utils.py:194-254)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import routing_np

pytestmark = pytest.mark.gpu

from vqa_model_builder_b200 import _lib, moe, ops  # noqa: E402
import vqa_model_builder_b200 as pkg  # noqa: E402

DEV = "cuda"


def ptr_array(tensors):
    return torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64, device=DEV)


@pytest.mark.parametrize("W,E,N,K", [(2, 8, 300, 2), (4, 8, 57, 2), (8, 8, 32, 2), (8, 32, 500, 2), (2, 4, 1, 1)])
def test_ep_layout_matches_numpy_oracle(W, E, N, K):
    rng = np.random.default_rng(W * 100 + E)
    idx = [np.stack([rng.permutation(E)[:K] for _ in range(N)]).astype(np.int32) for _ in range(W)]
    idx[0][rng.random(idx[0].shape) < 0.05] = -1                   # ablation-style masked pairs are never sent
    if W > 2:
        idx[1][:] = idx[1] % max(1, E // W)                          # rank 1 routes everything to rank 0's experts
    tab = np.stack([routing_np.routing_plan(i, E)["counts"] for i in idx])
    NK, El = N * K, E // W
    Rcap = _lib.query("b200_moe_max_rows", W * NK, El)
    tab_d = torch.from_numpy(tab.astype(np.int32)).to(DEV)
    for me in range(W):
        send_base = torch.empty(E, dtype=torch.int32, device=DEV)
        pad_off2 = torch.empty(2 * El + 1, dtype=torch.int32, device=DEV)
        tile_group2 = torch.empty(Rcap // 128, dtype=torch.int32, device=DEV)
        row_home = torch.empty(Rcap, dtype=torch.int32, device=DEV)
        _lib.ensure_device(tab_d)
        _lib.call("b200_ep_layout", tab_d, me, W, E, Rcap, NK, send_base, pad_off2, tile_group2, row_home,
                  _lib.stream_ptr())
        want = routing_np.ep_layout(tab, me, Rcap, NK)
        for name, got in (("send_base", send_base), ("pad_off2", pad_off2), ("tile_group2", tile_group2),
                          ("row_home", row_home)):
            assert np.array_equal(got.cpu().numpy(), want[name]), (name, me)


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
@pytest.mark.parametrize("W,E", [(2, 8), (4, 8), (8, 8)])
def test_emulated_expert_parallel_forward_backward_matches_unsharded_layer(mode, tol, W, E):
    torch.manual_seed(W + E)
    N, K, D, F = 96, 2, 128, 256
    El, NK = E // W, N * K
    pkg.set_compute_dtype(mode)
    try:
        cdt = torch.bfloat16 if mode == "bf16" else torch.float32
        full = moe.MOELayer(input_dim=D, hidden_dim=F, output_dim=D, num_experts=E, top_k=K, dropout=0.0).to(DEV).train()
        xs = [torch.randn(1, N, D, device=DEV).to(cdt).float() for _ in range(W)]
        gouts = [torch.randn(1, N, D, device=DEV) for _ in range(W)]
        # ---- unsharded layer on the concatenated batch --------------------------------------------------------------
        x_all = torch.cat(xs, dim=1).requires_grad_()
        ref = full(x_all)
        (ref * torch.cat(gouts, dim=1)).sum().backward()
        ref_dx = x_all.grad.view(W, N, D)
        ref_grads = {k: p.grad.clone() for k, p in full.named_parameters()}
        ref = ref.detach().view(W, N, D)
        for p in full.parameters():
            p.grad = None
        # ---- W emulated ranks ----------------------------------------------------------------------------------------
        Rcap = _lib.query("b200_moe_max_rows", W * NK, El)
        dt = _lib.dtype_code(cdt)
        sp = _lib.stream_ptr()
        i32 = dict(dtype=torch.int32, device=DEV)
        recv_x = [torch.full((Rcap, D), float("nan"), dtype=cdt, device=DEV) for _ in range(W)]
        recv_g = [torch.full((Rcap, D), float("nan"), dtype=cdt, device=DEV) for _ in range(W)]
        ret_z = [torch.zeros((NK, D), dtype=cdt, device=DEV) for _ in range(W)]
        ret_dx = [torch.zeros((NK, D), dtype=cdt, device=DEV) for _ in range(W)]
        tabs = [torch.zeros(W * E, **i32) for _ in range(W)]
        p_rx, p_rg, p_rz, p_rd, p_tab = map(ptr_array, (recv_x, recv_g, ret_z, ret_dx, tabs))
        x2s = [ops.to_compute(x.reshape(N, D), cdt).detach().requires_grad_() for x in xs]
        routed, plans, metas = [], [], []
        for r in range(W):                                    # phase 1: route, plan, push counts
            w, idx, aux = full.router(x2s[r].view(1, N, D))
            plan = ops.RoutingPlan(aux["_b200_idx32"][1], E, with_cmp_src=True)
            _lib.call("b200_ep_push_counts", plan.counts, p_tab, r, W, E, sp)
            routed.append(w)
            plans.append(plan)
        for r in range(W):                                    # phase 2: layout + dispatch into the owners' GEMM inputs
            m = dict(send_base=torch.empty(E, **i32), pad_off2=torch.empty(2 * El + 1, **i32),
                     tile_group2=torch.empty(Rcap // 128, **i32), row_home=torch.empty(Rcap, **i32))
            _lib.call("b200_ep_layout", tabs[r], r, W, E, Rcap, NK, m["send_base"], m["pad_off2"], m["tile_group2"],
                      m["row_home"], sp)
            _lib.call("b200_ep_dispatch", x2s[r].detach(), plans[r].cmp_src, plans[r].cmp_off, m["send_base"],
                      m["pad_off2"], p_rx, r, K, NK, E, El, D, Rcap, dt, sp)
            metas.append(m)
        assert all(torch.equal(t, tabs[0]) for t in tabs)
        slab = full._get_slab(torch.device(DEV), cdt)
        stacks = full._expert_stacks(torch.device(DEV), cdt)
        params = full._expert_params()                        # grouped per attribute: E entries each
        outs, z2s, xps = [], [], []
        for r in range(W):                                    # phase 3: local experts, return to the home ranks
            lo = r * El
            st_r = tuple(s[lo:lo + El] for s in stacks)
            pr = [p for a in range(6) for p in params[a * E + lo:a * E + lo + El]]
            used = int(metas[r]["pad_off2"][El])
            assert torch.isfinite(recv_x[r][:used].float()).all(), "a row in use was never written"
            xp = recv_x[r].detach().requires_grad_()
            z2 = ops.ExpertFFNFn.apply(xp, metas[r]["tile_group2"], metas[r]["pad_off2"], st_r, full.experts[0].act_code,
                                       True, full.experts[0].layer_norm.eps, None, *pr)
            _lib.call("b200_ep_return", z2.detach(), metas[r]["row_home"], metas[r]["pad_off2"], p_rz, El, D, Rcap, NK,
                      NK, dt, sp)
            z2s.append(z2)
            xps.append(xp)
        backs = []
        for r in range(W):                                    # combine at home
            back = ret_z[r].detach().requires_grad_()
            out = ops.CombineFn.apply(back, routed[r].reshape(N, K).float(), plans[r].cmp_pos, plans[r].cmp_src,
                                      full.output_norm.weight, full.output_norm.bias, full.output_norm.eps)
            outs.append(out)
            backs.append(back)
            assert rel_err(out, ref[r]) < tol, (r, rel_err(out, ref[r]))
        # ---- backward: the same two kernels in the opposite direction -------------------------------------------------
        for r in range(W):
            (outs[r].float() * gouts[r].view(N, D)).sum().backward()
            _lib.call("b200_ep_dispatch", backs[r].grad.to(cdt), None, plans[r].cmp_off, metas[r]["send_base"],
                      metas[r]["pad_off2"], p_rg, r, K, NK, E, El, D, Rcap, dt, sp)
        for r in range(W):
            used = int(metas[r]["pad_off2"][El])
            dz2 = recv_g[r].clone()
            dz2[used:] = 0
            z2s[r].backward(dz2)
            _lib.call("b200_ep_return", xps[r].grad.contiguous(), metas[r]["row_home"], metas[r]["pad_off2"], p_rd, El, D,
                      Rcap, NK, NK, dt, sp)
        for r in range(W):
            dx = torch.empty((N, D), dtype=cdt, device=DEV)
            _lib.call("b200_moe_unpermute", ret_dx[r], plans[r].cmp_pos, None, N, K, D, dt, dx, sp)
            total = dx.float() + x2s[r].grad.float()                      # expert path + router path
            assert rel_err(total, ref_dx[r]) < tol, (r, rel_err(total, ref_dx[r]))
        worst = max((rel_err(p.grad, ref_grads[k]), k) for k, p in full.named_parameters())
        assert worst[0] < tol, worst
    finally:
        pkg.set_compute_dtype("auto")
