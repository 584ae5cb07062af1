"""tcnn.Network / tcnn.NetworkWithInputEncoding shaped modules (supernormal_b200.tcnn_api; SURVEY.md 8b "optional", the interface
nerfacc's examples/radiance_fields/ngp.py:108-145 builds).  CPU: construction, parameter layout, conventions.  GPU: forward against a
plain fp32 expression of the same bias-free MLP (fp16 tolerance), autograd to params / inputs, one flat `params` in state_dict."""
import pytest
import torch

NET = {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "Sigmoid", "n_neurons": 64, "n_hidden_layers": 2}
ENC = {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 15, "base_resolution": 16, "per_level_scale": 1.4472692012786865}


def test_network_layout_and_conventions():
    from supernormal_b200 import tcnn_api as tcnn
    n = tcnn.Network(19, 3, NET)
    assert n.n_input_dims == 19 and n.n_output_dims == 3 and n.padded_input == 32 and n.padded_output == 16
    assert n.params.dtype == torch.float32 and n.params.numel() == 64 * 32 + 64 * 64 + 16 * 64 and list(n.state_dict()) == ["params"]
    assert torch.equal(n.params, tcnn.Network(19, 3, NET, seed=1337).params) and not torch.equal(n.params, tcnn.Network(19, 3, NET, seed=7).params)
    w = n._weights(n.params)
    assert [tuple(x.shape) for x in w] == [(64, 32), (64, 64), (16, 64)]
    assert all(x.abs().max() <= (6.0 / sum(x.shape)) ** 0.5 for x in w)          # xavier-uniform bound per matrix
    with pytest.raises(NotImplementedError):
        tcnn.Network(3, 1, dict(NET, activation="Snake"))
    with pytest.raises(NotImplementedError):
        n(torch.zeros(4, 19))                                                      # CUDA only, like the rest of the package
    m = tcnn.NetworkWithInputEncoding(3, 16, ENC, dict(NET, output_activation="None", n_hidden_layers=1))
    enc = tcnn.Encoding(3, ENC)
    assert list(m.state_dict()) == ["params"] and m.params.numel() == (64 * 32 + 16 * 64) + enc.params.numel()
    assert torch.equal(m.params[64 * 32 + 16 * 64:], enc.params)                   # [network | encoding]


@pytest.mark.gpu
def test_network_forward_backward_vs_fp32(cuda):
    from supernormal_b200 import tcnn_api as tcnn
    n = tcnn.Network(19, 3, NET).to(cuda)
    x = torch.rand(1000, 19, device=cuda, requires_grad=True)
    y = n(x)
    assert y.dtype == torch.float16 and y.shape == (1000, 3)
    w = [t.float() for t in n._weights(n.params)]
    h = torch.nn.functional.pad(x.detach(), (0, 13), value=1.0)
    h = torch.relu(h @ w[0].T)
    h = torch.relu(h @ w[1].T)
    ref = torch.sigmoid(h @ w[2].T)[:, :3]
    assert (y.float() - ref).abs().max().item() < 5e-3
    y.float().sum().backward()
    assert n.params.grad is not None and torch.isfinite(n.params.grad).all() and n.params.grad.abs().sum() > 0 and x.grad.abs().sum() > 0
    m = tcnn.NetworkWithInputEncoding(3, 16, ENC, dict(NET, output_activation="None", n_hidden_layers=1)).to(cuda)
    p = torch.rand(777, 3, device=cuda)
    out = m(p)
    assert out.shape == (777, 16) and out.dtype == torch.float16
    enc = tcnn.Encoding(3, ENC).to(cuda)
    feat = enc(p)
    assert torch.equal(out, m._net.mlp(feat, m.params[:m._n_net]))                 # encoding half == the stand-alone Encoding on the same table
    out.float().pow(2).sum().backward()
    g = m.params.grad
    assert torch.isfinite(g).all() and g[:m._n_net].abs().sum() > 0 and g[m._n_net:].abs().sum() > 0
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    opt.step()
    assert not torch.equal(m(p), out)                                              # the fp16 table cache follows the optimizer's in-place update
