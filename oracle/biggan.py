"""CPU restatement of the BigGAN-deep generator -- TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED.**

BASELINE.json's north_star / configs[4] name "BigGAN-deep 256 generator (GBlock: class-conditional BN + upsample +
3x3 conv + self-attention)", but /root/reference (alexandonian/pretorched-x @ 36a5754) contains **no GAN code at all**
(SURVEY.md section 0.4 / 8a row a14 / 8f row n1): there is no reference forward to run, no call site and no golden
vector to pin against.  This file therefore restates the *published* architecture -- "Large Scale GAN Training for High
Fidelity Natural Image Synthesis" (Brock et al. 2019), appendix B table 8/9, as implemented by the authors' public
`BigGANdeep.py` / `layers.py` (ajbrock/BigGAN-PyTorch, not vendored, not pinned by the reference) -- from the paper's
description, in plain functional torch on CPU fp32.  Every parity claim that rests on it is "our CUDA path equals our
own restatement", nothing stronger; DESIGN.md says the same.

Only ``tests/``, ``__graft_entry__`` and ``bench.py`` / ``tools/`` CPU legs may import this module.

Architecture (ch = 128 for the 256x256 model):
  y      = [shared_embedding(class) (128) | z (128)]                       (hier=True: the whole z goes to every block)
  h      = SNLinear(256 -> 16ch * 4*4)(y).view(B, 16ch, 4, 4)
  stage i (in -> out, attention after the 64x64 stage):
           GBlock(in, in), GBlock(in, out, nearest 2x upsample), [SAGAN Attention(out)]
  image  = tanh(SNConv3x3(ReLU(BN(h))))  ->  [B, 3, R, R]
GBlock(x, y) with hidden = in/4:
  h = conv1x1_1(relu(ccbn1(x, y)));  h = relu(ccbn2(h, y));  x = x[:, :out];  (h, x = upsample(h), upsample(x))
  h = conv3x3_2(h);  h = conv3x3_3(relu(ccbn3(h, y)));  h = conv1x1_4(relu(ccbn4(h, y)));  return h + x
ccbn(x, y) = batch_norm(x; stored_mean, stored_var, eps) * (1 + SNLinear_gain(y)) + SNLinear_bias(y)      (no biases)
Attention(x): theta = conv1x1(C -> C/8), phi = maxpool2(conv1x1(C -> C/8)), g = maxpool2(conv1x1(C -> C/2)),
              o = conv1x1(C/2 -> C)(g . softmax(theta^T phi)^T);  return gamma * o + x
Spectral norm (eval): W / sigma with sigma from ONE power-iteration step off the stored left vector u0.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
SN_EPS = 1e-12

# resolution -> (in multipliers, out multipliers, attention-after-stage flags) for ch-multiples
ARCH = {
    512: ([16, 16, 8, 8, 4, 2, 1], [16, 8, 8, 4, 2, 1, 1], {64}),
    256: ([16, 16, 8, 8, 4, 2], [16, 8, 8, 4, 2, 1], {64}),
    128: ([16, 16, 8, 4, 2], [16, 8, 4, 2, 1], {64}),
    64: ([16, 16, 8, 4], [16, 8, 4, 2], {64}),
    32: ([4, 4, 4], [4, 4, 4], set()),
}


def plan(resolution=256, ch=128):
    """[(in_channels, out_channels, output resolution, has_attention)] per stage."""
    ins, outs, att = ARCH[resolution]
    return [(ch * i, ch * o, 8 << k, (8 << k) in att) for k, (i, o) in enumerate(zip(ins, outs))]


def sn_sigma(w, u, eps=SN_EPS):
    """Largest singular value estimate of W (viewed [out, -1]) from one power iteration started at u [1, out]."""
    wm = w.reshape(w.shape[0], -1)
    v = F.normalize(torch.matmul(u, wm), eps=eps)
    u2 = F.normalize(torch.matmul(v, wm.t()), eps=eps)
    return torch.squeeze(torch.matmul(torch.matmul(v, wm.t()), u2.t()))


def _ident(t):
    return t


def fp16_storage(t):
    """Round-trip through fp16: what ``storage=fp16_storage`` applies at every point where the CUDA engine keeps a tensor
    in fp16 (weights, every activation it writes to HBM; accumulation, BN scale/shift and softmax stay fp32).  The
    restatement run this way is the *expected-numerics* twin of the product path: its distance from the plain fp32
    restatement is what fp16 storage costs by construction, and the GPU tests bound the product's error by it."""
    return t.half().float()


def sn_weight(sd, p, q=_ident):
    return q(sd[p + '.weight'] / sn_sigma(sd[p + '.weight'], sd[p + '.u0']))


def _sn_conv(x, sd, p, padding, q=_ident):
    return F.conv2d(x, sn_weight(sd, p, q), sd.get(p + '.bias'), 1, padding)


def ccbn(x, y, sd, p, eps=BN_EPS):
    gain = (1.0 + F.linear(y, sn_weight(sd, p + '.gain'))).view(y.size(0), -1, 1, 1)
    bias = F.linear(y, sn_weight(sd, p + '.bias')).view(y.size(0), -1, 1, 1)
    out = F.batch_norm(x, sd[p + '.stored_mean'], sd[p + '.stored_var'], None, None, False, 0.1, eps)
    return out * gain + bias


def gblock(x, y, sd, p, out_channels, upsample, q=_ident):
    h = _sn_conv(q(F.relu(ccbn(x, y, sd, p + '.bn1'))), sd, p + '.conv1', 0, q)
    h = q(F.relu(ccbn(h, y, sd, p + '.bn2')))
    if x.shape[1] != out_channels:
        x = x[:, :out_channels]
    if upsample:
        h = F.interpolate(h, scale_factor=2)          # nearest
        x = F.interpolate(x, scale_factor=2)
    h = _sn_conv(h, sd, p + '.conv2', 1, q)
    h = _sn_conv(q(F.relu(ccbn(h, y, sd, p + '.bn3'))), sd, p + '.conv3', 1, q)
    h = _sn_conv(q(F.relu(ccbn(h, y, sd, p + '.bn4'))), sd, p + '.conv4', 0, q)
    return q(h + x)


def attention(x, sd, p, q=_ident):
    B, C, H, W = x.shape
    theta = q(_sn_conv(x, sd, p + '.theta', 0, q))
    phi = F.max_pool2d(q(_sn_conv(x, sd, p + '.phi', 0, q)), [2, 2])
    g = F.max_pool2d(q(_sn_conv(x, sd, p + '.g', 0, q)), [2, 2])
    theta = theta.view(B, C // 8, H * W)
    phi = phi.view(B, C // 8, H * W // 4)
    g = g.view(B, C // 2, H * W // 4)
    beta = q(F.softmax(torch.bmm(theta.transpose(1, 2), phi), -1))
    o = _sn_conv(q(torch.bmm(g, beta.transpose(1, 2)).view(B, C // 2, H, W)), sd, p + '.o', 0, q)
    return q(sd[p + '.gamma'] * o + x)


def condition(z, labels, sd):
    """y = [shared(class) | z]  (hierarchical-z variant of BigGAN-deep: one vector for every block)."""
    return torch.cat([sd['shared.weight'][labels], z], 1)


def generator_forward(z, labels, sd, resolution=256, ch=128, stages=None, eps=BN_EPS, storage=None):
    """z fp32 [B, dim_z], labels int64 [B] -> images fp32 [B, 3, R, R] in (-1, 1).  ``stages``: optional dict that
    receives the activation after every stage (and the linear / pre-tanh tensors) for stage-wise parity checks.
    ``storage``: None = plain fp32 restatement; ``fp16_storage`` = the expected-numerics twin (see there)."""
    q = storage or _ident
    y = condition(z, labels, sd)
    pl = plan(resolution, ch)
    h = q(F.linear(q(y), sn_weight(sd, 'linear', q), sd['linear.bias'])).view(z.size(0), pl[0][0], 4, 4)
    if stages is not None:
        stages['linear'] = h
    for i, (cin, cout, res, att) in enumerate(pl):
        h = gblock(h, y, sd, 'blocks.%d.0' % i, cin, False, q)
        h = gblock(h, y, sd, 'blocks.%d.1' % i, cout, True, q)
        if att:
            h = attention(h, sd, 'blocks.%d.2' % i, q)
        if stages is not None:
            stages['stage%d' % i] = h
    p = 'output_layer'
    h = F.batch_norm(h, sd[p + '.0.stored_mean'], sd[p + '.0.stored_var'], sd[p + '.0.gain'], sd[p + '.0.bias'],
                     False, 0.1, eps)
    h = q(F.relu(h))
    if stages is not None:
        stages['out_act'] = h
    h = q(_sn_conv(h, sd, p + '.2', 1, q))
    if stages is not None:
        stages['pre_tanh'] = h
    return torch.tanh(h)


# ---------------------------------------------------------------------------------------------
# fixture conditioning: "standing statistics" + a non-trivial attention gate
# ---------------------------------------------------------------------------------------------
def calibrate_standing_stats_(sd, z, labels, resolution=256, ch=128, gamma=0.5, eps=BN_EPS):
    """Fill every stored_mean / stored_var of ``sd`` (in place) with the batch statistics a forward pass over (z, labels)
    sees at that layer -- what BigGAN's "standing statistics" evaluation mode does -- so that a random-init generator has
    O(1) activations at every depth (without it the 13 residual blocks drift by orders of magnitude and fp16-vs-fp32
    comparisons say nothing).  Also sets the attention gate gamma (initialised to 0 = attention disabled) to ``gamma``.
    Runs the restatement itself, layer by layer."""
    for k in sd:
        if k.endswith('.gamma'):
            sd[k] = torch.tensor(float(gamma))

    def stats_(x, p):
        sd[p + '.stored_mean'] = x.mean(dim=(0, 2, 3)).detach().clone()
        sd[p + '.stored_var'] = x.var(dim=(0, 2, 3), unbiased=False).detach().clone()

    y = condition(z, labels, sd)
    pl = plan(resolution, ch)
    h = F.linear(y, sn_weight(sd, 'linear'), sd['linear.bias']).view(z.size(0), pl[0][0], 4, 4)
    for i, (cin, cout, res, att) in enumerate(pl):
        for j, (oc, up) in enumerate(((cin, False), (cout, True))):
            p = 'blocks.%d.%d' % (i, j)
            x = h
            stats_(x, p + '.bn1')
            t = _sn_conv(F.relu(ccbn(x, y, sd, p + '.bn1')), sd, p + '.conv1', 0)
            stats_(t, p + '.bn2')
            t = F.relu(ccbn(t, y, sd, p + '.bn2'))
            if x.shape[1] != oc:
                x = x[:, :oc]
            if up:
                t = F.interpolate(t, scale_factor=2)
                x = F.interpolate(x, scale_factor=2)
            t = _sn_conv(t, sd, p + '.conv2', 1)
            stats_(t, p + '.bn3')
            t = _sn_conv(F.relu(ccbn(t, y, sd, p + '.bn3')), sd, p + '.conv3', 1)
            stats_(t, p + '.bn4')
            t = _sn_conv(F.relu(ccbn(t, y, sd, p + '.bn4')), sd, p + '.conv4', 0)
            h = t + x
        if att:
            h = attention(h, sd, 'blocks.%d.2' % i)
    stats_(h, 'output_layer.0')
    return sd


def mac_count(resolution=256, ch=128, dim_z=128, shared_dim=128):
    """Algorithmic multiply-accumulates per image (padding taps included, as hooking nn.Conv2d would count them)."""
    pl = plan(resolution, ch)
    cond = dim_z + shared_dim
    macs = cond * pl[0][0] * 16
    for cin, cout, res, att in pl:
        hid = cin // 4
        lo, hi = (res // 2) ** 2, res ** 2
        for oc, p_in, p_out in ((cin, lo, lo), (cout, lo, hi)):
            macs += cond * 2 * (cin + 3 * hid)                          # gain + bias linears of the four ccbn
            macs += p_in * cin * hid + 2 * p_out * hid * hid * 9 + p_out * hid * oc
        if att:
            C = cout
            macs += hi * C * (C // 8) * 2 + hi * C * (C // 2) + hi * (hi // 4) * (C // 8 + C // 2) + hi * (C // 2) * C
    macs += resolution ** 2 * pl[-1][1] * 3 * 9
    return macs


# ---------------------------------------------------------------------------------------------
# seeded test cases (shared by oracle/make_golden_biggan.py, tests/ and bench tools)
# ---------------------------------------------------------------------------------------------
def seeded_inputs(B, n_classes, seed, dim_z=128):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, dim_z, generator=g), torch.randint(0, n_classes, (B,), generator=g)


def build_case(factory, resolution, ch, n_classes, B, seed_init=0, seed_input=1, init='N02', gamma=0.5):
    """Seeded generator (``factory`` = pretorched_x_b200.biggan_deep) with calibrated standing statistics and an open
    attention gate.  Returns (model with the calibrated state loaded, state_dict, z, labels).  'N02' init draws
    torch.randn only (machine independent); 'ortho' goes through a QR whose rounding may differ between hosts."""
    torch.manual_seed(seed_init)
    model = factory(resolution, G_ch=ch, n_classes=n_classes, G_init=init)
    z, labels = seeded_inputs(B, n_classes, seed_input)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        calibrate_standing_stats_(sd, z, labels, resolution, ch, gamma=gamma)
    model.load_state_dict(sd)
    return model.eval(), sd, z, labels
