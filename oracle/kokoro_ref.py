"""CPU oracle: PyTorch fp32 restatement of the Kokoro-82M forward pass.

TEST INFRASTRUCTURE ONLY.  Nothing under ``kokorox_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs do, and there only as the checker / the timed CPU baseline.

PARITY UNPINNED: the reference (byteowlz/kokorox) contains no model arithmetic at all.  The hot
path is one opaque call, ``OrtKoko::infer`` (/root/reference/kokorox/src/onn/ort_koko.rs:37-91),
which hands ``input_ids`` [B,N] i64, ``style`` [B,256] f32 and ``speed`` [1] f32
(ort_koko.rs:56-75) to ONNX Runtime (crate ``ort`` 2.0.0-rc.11, Cargo.lock:2639-2661) executing
``kokoro-v1.0.onnx`` (koko.rs:57, hf_cache.rs:8-10) -- a third-party file that is absent from
/root/reference, as are onnxruntime and any golden audio.  No reference test constructs
``OrtKoko``.  This file therefore restates the *published* upstream algorithm
(hexgrad/Kokoro-82M ``kokoro/{model,modules,istftnet}.py``, summarised in SURVEY.md Appendix A);
what IS pinned by the reference and honoured here:

  * input contract and padding ``[0] + ids + [0]``      koko.rs:1168-1175, ort_koko.rs:56-75
  * token-id domain 0..177                                tts/vocab.rs:5-20, tokenize.rs:119-129
  * the 23-id example sequence                            ort_koko.rs:46
  * style vector = 256 floats; row = un-padded token count   koko.rs:1255-1306
  * 24 kHz mono f32 output, flattened                     koko.rs:59, :1179

Layout convention for dumped stages: time-major ``[L, C]`` (the CUDA path's layout), fp32.

Functions cite the SURVEY.md Appendix-A paragraph they restate.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

SAMPLE_RATE = 24000
N_HARM = 9            # harmonic_num 8 + fundamental  (A.9)
HOP_F0 = 300          # prod(upsample_rates) * gen_istft_hop_size = 10*6*5
N_FFT = 20
HOP = 5


def _t(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a
    a = np.ascontiguousarray(a)
    if not a.flags.writeable:
        a = a.copy()
    return torch.from_numpy(a)


class KokoroOracle:
    """Functional fp32 forward over a flat ``name -> tensor`` weight dict (weight-norm folded)."""

    def __init__(self, weights: Dict[str, np.ndarray], threads: Optional[int] = None,
                 stft_pad_mode: str = "reflect"):
        if threads:
            torch.set_num_threads(int(threads))
        self.w = {k: _t(v).float() for k, v in weights.items()}
        self.stft_pad_mode = stft_pad_mode
        self._lstm_cache: Dict[str, torch.nn.LSTM] = {}
        self.window = torch.hann_window(N_FFT, periodic=True, dtype=torch.float32)

    # ------------------------------------------------------------------ helpers
    def _lin(self, x, name, bias=True):
        return F.linear(x, self.w[name + ".weight"], self.w[name + ".bias"] if bias else None)

    def _conv(self, x, name, stride=1, padding=0, dilation=1, bias=True, groups=1):
        return F.conv1d(x, self.w[name + ".weight"], self.w[name + ".bias"] if bias else None,
                        stride=stride, padding=padding, dilation=dilation, groups=groups)

    def _lstm(self, prefix: str, x: torch.Tensor) -> torch.Tensor:
        """Bidirectional 1-layer LSTM, torch gate order i,f,g,o (A.11).  x [N,In] -> [N,2H]."""
        m = self._lstm_cache.get(prefix)
        if m is None:
            w_ih = self.w[prefix + ".weight_ih_l0"]
            H = w_ih.shape[0] // 4
            m = torch.nn.LSTM(w_ih.shape[1], H, 1, batch_first=True, bidirectional=True)
            sd = {k: self.w[f"{prefix}.{k}"] for k in (
                "weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0",
                "weight_ih_l0_reverse", "weight_hh_l0_reverse", "bias_ih_l0_reverse",
                "bias_hh_l0_reverse")}
            m.load_state_dict(sd)
            m.eval()
            self._lstm_cache[prefix] = m
        with torch.no_grad():
            y, _ = m(x.unsqueeze(0))
        return y[0]

    # ------------------------------------------------------------------ A.2 ALBERT
    def albert(self, ids: torch.Tensor) -> torch.Tensor:
        """A.2: shared-layer ALBERT applied 12x, post-LN, gelu_new.  ids [N] -> [N,768]."""
        w = self.w
        N = ids.shape[0]
        e = (w["bert.embeddings.word_embeddings.weight"][ids]
             + w["bert.embeddings.position_embeddings.weight"][:N]
             + w["bert.embeddings.token_type_embeddings.weight"][0])
        e = F.layer_norm(e, (128,), w["bert.embeddings.LayerNorm.weight"],
                         w["bert.embeddings.LayerNorm.bias"], 1e-12)
        h = self._lin(e, "bert.encoder.embedding_hidden_mapping_in")
        L = "bert.encoder.albert_layer_groups.0.albert_layers.0."
        for _ in range(12):
            q = self._lin(h, L + "attention.query").view(N, 12, 64).transpose(0, 1)
            k = self._lin(h, L + "attention.key").view(N, 12, 64).transpose(0, 1)
            v = self._lin(h, L + "attention.value").view(N, 12, 64).transpose(0, 1)
            s = torch.matmul(q, k.transpose(1, 2)) / 8.0
            p = torch.softmax(s, dim=-1)
            c = torch.matmul(p, v).transpose(0, 1).reshape(N, 768)
            a = self._lin(c, L + "attention.dense")
            h = F.layer_norm(h + a, (768,), w[L + "attention.LayerNorm.weight"],
                             w[L + "attention.LayerNorm.bias"], 1e-12)
            f = self._lin(h, L + "ffn")
            f = 0.5 * f * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (f + 0.044715 * f ** 3)))
            f = self._lin(f, L + "ffn_output")
            h = F.layer_norm(f + h, (768,), w[L + "full_layer_layer_norm.weight"],
                             w[L + "full_layer_layer_norm.bias"], 1e-12)
        return h

    # ------------------------------------------------------------------ A.3
    def duration_encoder(self, d_en: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
        """A.3: 3x (biLSTM -> AdaLayerNorm -> concat style).  d_en [N,512], s [128] -> [N,640]."""
        N = d_en.shape[0]
        sb = s.unsqueeze(0).expand(N, -1)
        x = torch.cat([d_en, sb], dim=1)
        for i in range(3):
            x = self._lstm(f"predictor.text_encoder.lstms.{2 * i}", x)
            h = self._lin(s, f"predictor.text_encoder.lstms.{2 * i + 1}.fc")
            gamma, beta = h[:512], h[512:]
            x = F.layer_norm(x, (512,), eps=1e-5)
            x = (1 + gamma) * x + beta
            x = torch.cat([x, sb], dim=1)
        return x

    # ------------------------------------------------------------------ A.4
    def text_encoder(self, ids: torch.Tensor) -> torch.Tensor:
        """A.4: embedding -> 3x(conv k5 -> channel LN -> LReLU 0.2) -> biLSTM.  -> [N,512]."""
        w = self.w
        x = w["text_encoder.embedding.weight"][ids].t().unsqueeze(0)  # [1,512,N]
        for i in range(3):
            x = self._conv(x, f"text_encoder.cnn.{i}.0", padding=2)
            x = F.layer_norm(x.transpose(1, 2), (512,), w[f"text_encoder.cnn.{i}.1.gamma"],
                             w[f"text_encoder.cnn.{i}.1.beta"], 1e-5).transpose(1, 2)
            x = F.leaky_relu(x, 0.2)
        return self._lstm("text_encoder.lstm", x[0].t())

    # ------------------------------------------------------------------ A.5 / A.6
    def _adain(self, x, s, name):
        """A.5: (1+gamma) * InstanceNorm1d(x; eps 1e-5, biased var) + beta.  x [1,C,L]."""
        h = self._lin(s, name + ".fc")
        C = x.shape[1]
        gamma, beta = h[:C].view(1, C, 1), h[C:].view(1, C, 1)
        return (1 + gamma) * F.instance_norm(x, eps=1e-5) + beta

    def _adain_resblk(self, x, s, name, upsample=False):
        """A.6 AdainResBlk1d.  x [1,Ci,L] -> [1,Co,L or 2L]."""
        w = self.w
        Ci = x.shape[1]
        r = self._adain(x, s, name + ".norm1")
        r = F.leaky_relu(r, 0.2)
        if upsample:
            r = F.conv_transpose1d(r, w[name + ".pool.weight"], w[name + ".pool.bias"], stride=2,
                                   padding=1, output_padding=1, groups=Ci)
        r = self._conv(r, name + ".conv1", padding=1)
        r = self._adain(r, s, name + ".norm2")
        r = F.leaky_relu(r, 0.2)
        r = self._conv(r, name + ".conv2", padding=1)
        sc = x
        if upsample:
            sc = F.interpolate(sc, scale_factor=2, mode="nearest")
        if (name + ".conv1x1.weight") in w:
            sc = self._conv(sc, name + ".conv1x1", bias=False)
        return (r + sc) * torch.rsqrt(torch.tensor(2.0))

    # ------------------------------------------------------------------ A.7
    def f0n(self, en: torch.Tensor, s: torch.Tensor):
        """A.7 F0Ntrain.  en [T,640] -> F0 [2T], N [2T], plus the shared-LSTM output [T,512]."""
        x = self._lstm("predictor.shared", en)          # [T,512]
        outs = []
        for br in ("F0", "N"):
            y = x.t().unsqueeze(0)
            y = self._adain_resblk(y, s, f"predictor.{br}.0")
            y = self._adain_resblk(y, s, f"predictor.{br}.1", upsample=True)
            y = self._adain_resblk(y, s, f"predictor.{br}.2")
            y = self._conv(y, f"predictor.{br}_proj")
            outs.append(y[0, 0])
        return outs[0], outs[1], x

    # ------------------------------------------------------------------ A.9 source
    def harmonic_source(self, f0: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        """A.9 SineGen + SourceModuleHnNSF.  f0 [2T], noise [600T,9] -> har_source [600T].

        rand_ini (upstream adds a random phase at time index 0 only) is omitted: the x1/300 linear
        down-sampling reads taps 300i+149 / 300i+150 only, so index 0 never reaches the output.
        """
        L2 = f0.shape[0]
        f0_up = f0.repeat_interleave(HOP_F0).view(1, -1, 1)                 # nearest x300
        harm = torch.arange(1, N_HARM + 1, dtype=torch.float32).view(1, 1, -1)
        fn = f0_up * harm
        rad = (fn / SAMPLE_RATE) % 1
        rad = F.interpolate(rad.transpose(1, 2), scale_factor=1 / HOP_F0, mode="linear").transpose(1, 2)
        assert rad.shape[1] == L2
        phase = torch.cumsum(rad, dim=1) * 2 * np.pi
        phase = F.interpolate(phase.transpose(1, 2) * HOP_F0, scale_factor=HOP_F0,
                              mode="linear").transpose(1, 2)
        sines = torch.sin(phase) * 0.1
        uv = (f0_up > 10).float()
        noise_amp = uv * 0.003 + (1 - uv) * 0.1 / 3
        src = sines * uv + noise_amp * noise.view(1, -1, N_HARM)
        har = torch.tanh(self._lin(src, "decoder.generator.m_source.l_linear"))
        return har[0, :, 0]

    def stft(self, x: torch.Tensor):
        """A.9 STFT: n_fft 20, hop 5, periodic hann, center.  x [S] -> mag, phase [11, S/5+1].

        Branch-cut canonicalisation: angle() is discontinuous at (re<0, im=0).  Frame 0 is
        even-symmetric about its centre (reflect padding about sample 0 + symmetric window), so every
        bin has im == 0 up to rounding and upstream's +pi / -pi there is decided by FFT rounding
        noise.  Both this oracle and the CUDA kernel map (re<0, |im| <= 4e-6 * sum_j |x_j w_j|), i.e.
        an imaginary part inside the fp32 rounding floor of the 20-term sum, to +pi;
        e^{i pi} = e^{-i pi}, so both choices denote the same spectrum."""
        X = torch.stft(x.unsqueeze(0), N_FFT, HOP, N_FFT, window=self.window, center=True,
                       pad_mode=self.stft_pad_mode, return_complex=True)[0]
        xp = F.pad(x.view(1, 1, -1), (N_FFT // 2, N_FFT // 2), mode=self.stft_pad_mode)[0, 0]
        l1 = (xp.unfold(0, N_FFT, HOP) * self.window).abs().sum(dim=1)          # [frames]
        ph = torch.angle(X)
        cut = (X.real < 0) & (X.imag.abs() <= 4e-6 * l1.unsqueeze(0))
        ph = torch.where(cut, torch.full_like(ph, math.pi), ph)
        return torch.abs(X), ph

    def istft(self, mag: torch.Tensor, ph: torch.Tensor) -> torch.Tensor:
        """A.9 head: torch.istft semantics (window OLA / envelope, trim n_fft/2)."""
        X = mag * torch.exp(ph * 1j)
        return torch.istft(X.unsqueeze(0), N_FFT, HOP, N_FFT, window=self.window)[0]

    def _adain_resblock1(self, x, s, name, k):
        """A.9 AdaINResBlock1(C,k) with Snake activations, dilations (1,3,5)."""
        w = self.w
        for j, d in enumerate((1, 3, 5)):
            a1 = w[f"{name}.alpha1.{j}"]
            a2 = w[f"{name}.alpha2.{j}"]
            xt = self._adain(x, s, f"{name}.adain1.{j}")
            xt = xt + (1 / a1) * (torch.sin(a1 * xt) ** 2)
            xt = self._conv(xt, f"{name}.convs1.{j}", padding=d * (k - 1) // 2, dilation=d)
            xt = self._adain(xt, s, f"{name}.adain2.{j}")
            xt = xt + (1 / a2) * (torch.sin(a2 * xt) ** 2)
            xt = self._conv(xt, f"{name}.convs2.{j}", padding=(k - 1) // 2)
            x = xt + x
        return x

    def generator(self, x, s, f0, noise, st):
        """A.9 Generator trunk + head.  x [1,512,2T] -> audio [600T]."""
        w = self.w
        G = "decoder.generator."
        har_src = self.harmonic_source(f0, noise)
        st["har_source"] = har_src
        mag, ph = self.stft(har_src)
        har = torch.cat([mag, ph], dim=0).unsqueeze(0)            # [1,22,120T+1]
        st["har"] = har[0].t()
        ups = ((10, 20), (6, 12))
        for i in range(2):
            x = F.leaky_relu(x, 0.1)
            if i == 0:
                xs = self._conv(har, G + "noise_convs.0", stride=6, padding=3)
                xs = self._adain_resblock1(xs, s, G + "noise_res.0", 7)
            else:
                xs = self._conv(har, G + "noise_convs.1")
                xs = self._adain_resblock1(xs, s, G + "noise_res.1", 11)
            st[f"gen.x_source.{i}"] = xs[0].t()
            u, k = ups[i]
            x = F.conv_transpose1d(x, w[G + f"ups.{i}.weight"], w[G + f"ups.{i}.bias"], stride=u,
                                   padding=(k - u) // 2)
            if i == 1:
                x = F.pad(x, (1, 0), mode="reflect")
            st[f"gen.ups.{i}"] = x[0].t()
            x = x + xs
            acc = None
            for j, kk in enumerate((3, 7, 11)):
                y = self._adain_resblock1(x, s, G + f"resblocks.{i * 3 + j}", kk)
                acc = y if acc is None else acc + y
            x = acc / 3
            st[f"gen.stage.{i}"] = x[0].t()
        x = F.leaky_relu(x, 0.01)
        x = self._conv(x, G + "conv_post", padding=3)
        st["conv_post"] = x[0].t()
        mag = torch.exp(x[0, :11])
        ph = torch.sin(x[0, 11:])
        return self.istft(mag, ph)

    # ------------------------------------------------------------------ A.8
    def decoder(self, asr, f0, n, s, noise, st):
        """A.8 Decoder.  asr [T,512], f0/n [2T], s [128] -> audio [600T]."""
        D = "decoder."
        a = asr.t().unsqueeze(0)
        f0d = self._conv(f0.view(1, 1, -1), D + "F0_conv", stride=2, padding=1)
        nd = self._conv(n.view(1, 1, -1), D + "N_conv", stride=2, padding=1)
        x = torch.cat([a, f0d, nd], dim=1)
        x = self._adain_resblk(x, s, D + "encode")
        st["dec.encode"] = x[0].t()
        asr_res = self._conv(a, D + "asr_res.0")
        for i in range(4):
            x = torch.cat([x, asr_res, f0d, nd], dim=1)
            x = self._adain_resblk(x, s, D + f"decode.{i}", upsample=(i == 3))
            st[f"dec.decode.{i}"] = x[0].t()
        return self.generator(x, s, f0, noise, st)

    # ------------------------------------------------------------------ A.1 top level
    @torch.no_grad()
    def forward(self, tokens, style, speed: float = 1.0, noise: Optional[np.ndarray] = None,
                noise_seed: int = 0, stages: bool = False, inject: Optional[dict] = None):
        """A.1 forward_with_tokens at B=1.

        tokens: int64 [N] incl. the leading/trailing 0 pads (koko.rs:1168-1173).
        style : float32 [256]; [:128] -> decoder, [128:] -> predictor (A.1).
        noise : float32 flat, >= 600*T*9 elements, element (sample t, harmonic h) at t*9+h;
                None -> N(0,1) from ``noise_seed``.
        inject: optional {"pred_dur": int array, "F0": [2T], "N": [2T]} teacher-forcing overrides
                (used by the stage-wise parity tests).
        Returns dict(audio [600T] f32, pred_dur [N] i64, stages {name: tensor}).
        """
        inject = inject or {}
        ids = _t(np.asarray(tokens, dtype=np.int64))
        style = _t(np.asarray(style, dtype=np.float32)).view(-1)
        assert style.numel() == 256
        s_dec, s_pro = style[:128], style[128:]
        N = ids.shape[0]
        st: Dict[str, torch.Tensor] = {}

        bert = self.albert(ids)
        st["bert"] = bert
        d_en = self._lin(bert, "bert_encoder")                          # [N,512]
        st["d_en"] = d_en
        d = self.duration_encoder(d_en, s_pro)                           # [N,640]
        st["d"] = d
        x = self._lstm("predictor.lstm", d)                              # [N,512]
        st["dur_lstm"] = x
        logits = self._lin(x, "predictor.duration_proj.linear_layer")    # [N,50]
        st["dur_logits"] = logits
        dur = torch.sigmoid(logits).sum(dim=-1) / speed
        st["dur_float"] = dur
        pred_dur = torch.round(dur).clamp(min=1).long()
        if "pred_dur" in inject:
            pred_dur = _t(np.asarray(inject["pred_dur"], dtype=np.int64))
        idx = torch.repeat_interleave(torch.arange(N), pred_dur)
        T = int(idx.shape[0])
        st["idx"] = idx
        en = d[idx]                                                      # [T,640]
        f0, n, shared = self.f0n(en, s_pro)
        st["shared_lstm"] = shared
        if "F0" in inject:
            f0 = _t(np.asarray(inject["F0"], dtype=np.float32))
        if "N" in inject:
            n = _t(np.asarray(inject["N"], dtype=np.float32))
        st["F0"], st["N"] = f0, n
        t_en = self.text_encoder(ids)                                    # [N,512]
        st["t_en"] = t_en
        asr = t_en[idx]
        if noise is None:
            g = torch.Generator().manual_seed(int(noise_seed))
            nz = torch.randn(600 * T * N_HARM, generator=g)
        else:
            nz = _t(np.asarray(noise, dtype=np.float32)).view(-1)[: 600 * T * N_HARM]
            assert nz.numel() == 600 * T * N_HARM, "noise buffer too small"
        audio = self.decoder(asr, f0, n, s_dec, nz.view(-1, N_HARM), st)
        assert audio.shape[0] == 600 * T
        out = {"audio": audio.numpy().copy(), "pred_dur": pred_dur.numpy().copy(), "T": T}
        if stages:
            out["stages"] = {k: v.detach().contiguous().numpy().copy() for k, v in st.items()}
        return out
