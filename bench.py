#!/usr/bin/env python
"""Benchmark of the wav2vec-S encoder forward path (BASELINE.json metric: audio-seconds encoded per
second, wav2vec-S large, bf16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path (`extract_features`) over one batch of synthetic waveforms.
Workloads (BASELINE.json configs): large_64x20s (default; configs[2], the largest single-GPU
config of the model the metric is quoted on), base_1x10s (configs[0], with --dtype fp32), base_32x15s
(configs[1]), stream_large_b1 / _b16 (configs[3]; also probed by the default run), large_64x30s (configs[4];
at N > 1 the default run adds it as `extra_points`).
With N > 1 (torchrun, one rank per GPU) every rank encodes its own batch -- utterances are
independent, no collective sits in the data path ("scaling": "weak"); NCCL only carries the
barrier and the max-over-ranks timing.

One JSON line is printed by rank 0.  `value` = device-resident throughput, `e2e` = the same metric
through the public API with pinned-host input and device->host read of the result inside the
timed region, `roofline` = the tcgen05 GEMM kernel (all launches of a step) against the measured
bf16 peak, `cpu_baseline` = the oracle port (reference algorithm in PyTorch) on the host cores.
`--impl reference` times that CPU implementation alone on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, B per GPU, seconds)
    "large_64x20s": ("large", 64, 20),
    "base_1x10s": ("base", 1, 10),                     # configs[0] (run with --dtype fp32: the reference's CPU-runnable case)
    "base_32x15s": ("base", 32, 15),
    "base_posconv_32x15s": ("base_posconv", 32, 15),   # same model with the wav2vec 2.0 convolutional positions
    "large_64x30s": ("large", 64, 30),
    "large_8x20s": ("large", 8, 20),
    "tiny_4x2s": ("tiny", 4, 2),
    # incremental (SimulEval-style) inference: configs[3]; B streams in lock-step, 30 s each
    "stream_large_b1": ("large", 1, 30),
    "stream_large_b16": ("large", 16, 30),
    "stream_tiny_b2": ("tiny", 2, 6),
}
SR = 16000


def model_cfg(kind):
    if kind == "large":
        return dict(extractor_mode="layer_norm", encoder_layers=24, encoder_embed_dim=1024,
                    encoder_ffn_embed_dim=4096, encoder_attention_heads=16, layer_norm_first=True,
                    conv_bias=True, pos_type="sin", main_context=16, right_context=8)
    if kind == "base_posconv":
        return dict(model_cfg("base"), pos_type="conv", conv_pos=128, conv_pos_groups=16)
    if kind == "base":
        return dict(extractor_mode="layer_norm", encoder_layers=12, encoder_embed_dim=768,
                    encoder_ffn_embed_dim=3072, encoder_attention_heads=12, layer_norm_first=False,
                    conv_bias=False, pos_type="sin", main_context=16, right_context=8)
    return dict(extractor_mode="layer_norm", encoder_layers=3, encoder_embed_dim=128,
                encoder_ffn_embed_dim=256, encoder_attention_heads=2, layer_norm_first=True,
                conv_bias=True, pos_type="sin", main_context=16, right_context=8,
                conv_feature_layers="[(64,10,5)] + [(64,3,2)]*4 + [(64,2,2)]*2")


def conv_spec(cfg):
    s = cfg.get("conv_feature_layers", "[(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512,2,2)] + [(512,2,2)]")
    return list(eval(s))


def flops_per_utt(cfg, L):
    """Algorithmic FLOPs (2*MAC, mask-aware attention) of one utterance; SURVEY.md section 8(d)."""
    spec = conv_spec(cfg)
    t, cin = L, 1
    conv = 0.0
    lens = []
    for (c, k, s) in spec:
        t = (t - k) // s + 1
        lens.append(t)
        conv += 2.0 * cin * c * k * t
        cin = c
    T = lens[-1]
    D, F, Ly = cfg["encoder_embed_dim"], cfg["encoder_ffn_embed_dim"], cfg["encoder_layers"]
    main, rc = cfg["main_context"], cfg["right_context"]
    T2 = T + (-T) % 2
    nb = T2 // main
    M = T2 + nb * rc
    P = 0
    for tt in range(T2):
        b = tt // main
        P += min(main * (b + 1), T2) + (rc if b < nb else 0)
    for b in range(nb):
        P += rc * (min(main * (b + 1), T2) + rc)
    proj = 2.0 * cin * D * T if cin != D else 0.0
    linear = Ly * M * (8.0 * D * D + 4.0 * D * F)
    attn = Ly * 4.0 * D * P
    posconv = 0.0
    if cfg.get("pos_type") == "conv":   # grouped Conv1d(D, D, k, groups), SURVEY.md section 8(d)
        posconv = 2.0 * D * (D // cfg["conv_pos_groups"]) * cfg["conv_pos"] * T
    gemm = (conv - 2.0 * spec[0][0] * spec[0][1] * lens[0]) + proj + linear + posconv   # what the GEMM kernel executes
    return dict(total=conv + proj + linear + attn + posconv, conv=conv, proj=proj, linear=linear, attn=attn,
                posconv=posconv, gemm=gemm, T=T, M=M)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def reference_cpu_encoder(cfg_kind):
    """(encode(wav) callable, kind, description): the reference's OWN implementation on the host CPU -- its
    unmodified modules through oracle/ref_shim.py (sources in the build container, oracle/_ref bytecode on the GPU
    box; kind "reference") -- or, when neither is present, the oracle port (kind "port")."""
    import warnings
    import torch
    from oracle import ref_shim, synth
    from oracle import w2vs_oracle as O
    warnings.filterwarnings("ignore")
    cfg = O.default_cfg(**model_cfg(cfg_kind))
    sd = synth.make_state_dict(cfg, 0)
    if ref_shim.available():
        m = ref_shim.build_fairseq_model(cfg)
        m.load_state_dict(sd, strict=False)

        def encode(wav):
            with torch.no_grad():
                return m.extract_features(wav, None)[0]
        return encode, "reference", f"unmodified reference modules ({ref_shim.kind()}) on torch CPU"
    return (lambda wav: O.extract_features(sd, cfg, wav, None)[0]), "port", "oracle port on torch CPU"


def cpu_reference_run(cfg_kind, B, seconds, steps, warmup, bf16_too=False):
    """Time the reference's CPU implementation (fp32, all host threads) on a bounded sample of the workload."""
    import torch
    from oracle import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    encode, kind, what = reference_cpu_encoder(cfg_kind)
    wav = synth.make_waveform(B, seconds * SR, 1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        encode(wav)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t = statistics.median(times)
    out = dict(value=B * seconds / t, unit="audio-s/s", cores=cores, kind=kind,
               sample=f"{cfg_kind} fp32, {B} x {seconds} s per step (a sample of the workload's batch), {what}, "
                      f"median of {steps} after {warmup} warm-up",
               ms_per_step=t * 1e3)
    if bf16_too and kind == "reference":
        # informational: the same reference modules in bf16 on the CPU (`model.bfloat16()`, bf16 waveform) -- faster
        # than its fp32 run on AMX/AVX512-bf16 hosts, and 2-2.6e-2 away from its own fp32 output
        try:
            from oracle import ref_shim
            from oracle import w2vs_oracle as O
            cfg = O.default_cfg(**model_cfg(cfg_kind))
            m16 = ref_shim.build_fairseq_model(cfg)
            m16.load_state_dict(synth.make_state_dict(cfg, 0), strict=False)
            m16 = m16.bfloat16()
            w16 = wav.to(torch.bfloat16)
            tt = []
            for i in range(2):
                t0 = time.perf_counter()
                with torch.no_grad():
                    m16.extract_features(w16, None)
                tt.append(time.perf_counter() - t0)
            out["bf16_value"] = B * seconds / min(tt)
        except Exception as ex:
            out["bf16_value"] = None
            out["bf16_note"] = f"{type(ex).__name__}: {ex}"
    return out


def stream_bench(a, kind, B, seconds):
    """configs[3]: chunk-by-chunk incremental inference with cached left context.  First chunk = 24 frames
    (7760 samples), then `--step-blocks` x 16 frames per decision step (5120 samples per block), as the
    SimulEval agent feeds the encoder (rain/simul/transducer_searcher.py:712-721).  Reports p50 per-chunk
    device latency; the reference arm re-encodes the whole prefix per step (what the reference does)."""
    import torch
    cfg = model_cfg(kind)
    L = seconds * SR
    step = 5120 * a.step_blocks
    bounds = [7760]
    while bounds[-1] + step < L:
        bounds.append(bounds[-1] + step)
    bounds.append(L)
    metric = f"p50 per-chunk latency (wav2vec-S {kind} incremental, batch {B}, {16 * a.step_blocks} frames/step)"
    config = {"workload": f"wav2vec-S {kind} incremental encoder, {B} stream(s) x {seconds} s, first chunk 24 frames, "
                          f"then {16 * a.step_blocks} frames per step, {a.dtype}"
                          + {0: "", 1: ", operator chain", 2: ", first persistent step kernel", 3: ""}[a.step_impl],
              "step_impl": a.step_impl, "name": a.workload,
              "chunks": len(bounds), "l2_policy": "weights (613 MB bf16) exceed the 126 MB L2"}
    if a.impl == "reference":
        import warnings
        from oracle import ref_shim, synth
        from oracle import w2vs_oracle as O
        warnings.filterwarnings("ignore")
        torch.set_num_threads(os.cpu_count() or 1)
        ocfg = O.default_cfg(**cfg)
        sd = synth.make_state_dict(ocfg, 0)
        wav = synth.make_waveform(1, L, 1234)
        sel = bounds[:: max(1, len(bounds) // 8)][:8]     # bounded sample of the decision steps
        ref_kind = "reference" if ref_shim.available() else "port"
        if ref_kind == "reference":       # the reference's own BlockWiseWav2Vec2Model, prefix re-encoding as its driver does
            rm = ref_shim.build_rain_model(ocfg)
            rm.load_state_dict(sd, strict=False)
        times = []
        for n in sel:
            t0 = time.perf_counter()
            if ref_kind == "reference":
                with torch.no_grad():
                    rm(wav[:, :n].clone(), None, None, n >= L, True)
            else:
                O.rain_forward(sd, ocfg, wav[:, :n], None, finished=(n >= L), is_infer=True)
            times.append((time.perf_counter() - t0) * 1e3)
        v = statistics.median(times)
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "ms", "n_gpus": a.gpus,
                          "steps": len(sel), "warmup": 0, "ms_per_step": v, "higher_is_better": False,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": v, "unit": "ms", "cores": os.cpu_count(), "kind": ref_kind,
                                           "sample": f"prefix re-encoding of 1 stream at {len(sel)} of {len(bounds)} decision steps"},
                          "e2e": {"value": v, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return
    import wav2vec_s_b200 as W
    from wav2vec_s_b200 import cabi
    from wav2vec_s_b200.model import EncoderStream
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    torch.manual_seed(0)
    model = W.BlockWiseWav2Vec2Model(cfg).to(dev, dtype).eval()
    g = torch.Generator().manual_seed(1234)
    wav_host = torch.randn(B, L, generator=g).pin_memory()
    wav = wav_host.to(dev)

    def run(e2e):
        st = model.open_stream(B=B, max_seconds=seconds + 1, max_new_samples=max(7760, step) + 400, step_impl=a.step_impl)
        lat, pos, outs = [], 0, 0
        for n in bounds:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record()
            chunk = wav_host[:, pos:n].to(dev, non_blocking=True) if e2e else wav[:, pos:n]
            y = st.step(chunk, EncoderStream.FINAL if n >= L else EncoderStream.NONE)
            if e2e:
                y = y.to("cpu", non_blocking=True)
            ev1.record()
            torch.cuda.synchronize()
            lat.append(ev0.elapsed_time(ev1))
            outs += y.size(0)
            pos = n
        return lat, outs

    for _ in range(max(1, a.warmup // 3)):
        run(False)
    cabi.launch_count(reset=True)
    sampler = ClockSampler(0)
    sampler.start()
    lats = []
    for _ in range(max(1, a.steps // 5)):
        lat, frames = run(False)
        lats += lat[1:-1]
    launches = cabi.launch_count(reset=True)
    clocks = sampler.stop()
    lat_e2e, _ = run(True)
    p50 = statistics.median(lats)
    fl = flops_per_utt(cfg, L)
    wbytes = 613e6 if kind == "large" else 179e6
    pk, pk_kind = peaks()
    line = {"metric": metric, "value": p50, "unit": "ms", "n_gpus": 1, "steps": len(lats), "warmup": a.warmup,
            "ms_per_step": p50, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype,
            "data": "synthetic", "config": config, "clocks": clocks,
            "p90_ms": sorted(lats)[int(0.9 * len(lats))], "frames_emitted": frames,
            "realtime_factor": (16 * a.step_blocks * 0.02) / (p50 / 1e3),
            "e2e": {"value": statistics.median(lat_e2e[1:-1]), "unit": "ms", "h2d_bytes_per_step": B * step * 4,
                    "d2h_bytes_per_step": 16 * a.step_blocks * B * cfg["encoder_embed_dim"] * 2},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "one decision step = weights streamed once per block",
                         "achieved": wbytes * a.step_blocks / (p50 / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": wbytes * a.step_blocks / (p50 / 1e3) / 1e9 / pk["hbm_gbs"], "traffic": None,
                         "peak_source": pk_kind}}
    print(json.dumps(line))


def incremental_probe(model, cfg, dev, B=1, seconds=30):
    """configs[3]: p50 / p90 device latency of a decision step (first chunk 24 frames, then 16 frames = 5120 samples per
    step) of B lock-step streams through BlockWiseWav2Vec2Model.open_stream -> w2vs_stream_step, on the bench's own
    weights, with its HBM roofline: a step has to read the packed bf16 weights once (613 MB for the large model) plus
    the cached K/V of the left context."""
    import torch
    import wav2vec_s_b200 as W
    from wav2vec_s_b200.model import EncoderStream
    sm = W.BlockWiseWav2Vec2Model(cfg)
    sm.load_state_dict(model.state_dict(), strict=False)
    sm = sm.to(dev, next(model.parameters()).dtype).eval()
    L = seconds * SR
    wav = torch.randn(B, L, generator=torch.Generator().manual_seed(4321)).to(dev)
    bounds = [7760]
    while bounds[-1] + 5120 < L:
        bounds.append(bounds[-1] + 5120)
    bounds.append(L)
    from wav2vec_s_b200 import cabi
    lats, launches = [], 0
    for rep in range(3):
        cabi.launch_count(reset=True)
        st = sm.open_stream(B=B, max_seconds=seconds + 1, max_new_samples=7760 + 400)
        pos, lat = 0, []
        for n in bounds:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            st.step(wav[:, pos:n], EncoderStream.FINAL if n >= L else EncoderStream.NONE)
            e1.record()
            torch.cuda.synchronize()
            lat.append(e0.elapsed_time(e1))
            pos = n
        launches = cabi.launch_count(reset=True) / len(bounds)
        if rep > 0:
            lats += lat[1:-1]
    lats.sort()
    p50 = lats[len(lats) // 2]
    D, Ly, F = cfg["encoder_embed_dim"], cfg["encoder_layers"], cfg["encoder_ffn_embed_dim"]
    wbytes = Ly * (4 * D * D + 2 * D * F) * 2 + (4 * 512 * 512 * 3 + 2 * 512 * 512 * 2 + 512 * D) * 2   # GEMM operands, bf16
    kvbytes = B * Ly * 2 * D * 2 * (seconds * 50 // 2)                 # cached K/V read at the mean left context
    pk, pk_kind = peaks()
    ach = (wbytes + kvbytes) / (p50 / 1e3) / 1e9
    return {"metric": f"p50 per-chunk latency (wav2vec-S large incremental, {B} stream(s), 16 frames/step)",
            "streams": B, "launches_per_step": round(launches, 1),
            "path": "one kernel of thread-block clusters per step (k_stream_cluster.cu) behind the conv-stack launches"
                    if launches < 60 else "kernel-per-operator chain with programmatic dependent launches",
            "p50_ms": p50, "p90_ms": lats[int(0.9 * len(lats))], "unit": "ms", "steps": len(lats),
            "realtime_factor": 0.32 / (p50 / 1e3),
            "roofline": {"bound": "hbm", "kernel": "one decision step (weights streamed once + cached K/V)",
                         "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                         "traffic": None, "algorithmic_bytes": wbytes + kvbytes, "peak_source": pk_kind}}


def gemm_bytes_per_utt(cfg, L):
    """Algorithmic HBM bytes of the GEMM launches of one utterance (bf16 model), weights not included: conv inputs and
    outputs, QKV / FFN operands and results, the fp32 residual stream read and written by out_proj and fc2."""
    spec = conv_spec(cfg)
    t, lens = L, []
    for (c, k, s_) in spec:
        t = (t - k) // s_ + 1
        lens.append(t)
    fl = flops_per_utt(cfg, L)
    T, M, D, F, Ly = fl["T"], fl["M"], cfg["encoder_embed_dim"], cfg["encoder_ffn_embed_dim"], cfg["encoder_layers"]
    conv = sum(2 * spec[i - 1][0] * lens[i - 1] + 2 * spec[i][0] * lens[i] for i in range(1, len(spec)))
    proj = 2 * spec[-1][0] * T + 4 * D * T
    layer = (2 * M * D + 2 * M * 3 * D) + (2 * M * D + 8 * M * D) + (2 * M * D + 2 * M * F) + (2 * M * F + 8 * M * D)
    return conv + proj + Ly * layer


def rows_bytes_per_utt(cfg, L):
    """Algorithmic HBM bytes of the stand-alone row passes of one utterance (bf16 model): conv LayerNorm + GELU passes
    (read + write bf16), feature LayerNorm, token embedding, two LayerNorms per transformer layer (read fp32, write
    bf16), final LayerNorm."""
    spec = conv_spec(cfg)
    t, lens = L, []
    for (c, k, s_) in spec:
        t = (t - k) // s_ + 1
        lens.append(t)
    n_ln = 1 if cfg["encoder_layers"] == 12 else 7
    conv_ln = sum(4 * spec[i][0] * lens[i] for i in range(1, len(spec)) if i < n_ln)
    fl = flops_per_utt(cfg, L)
    T, M, D, Ly = fl["T"], fl["M"], cfg["encoder_embed_dim"], cfg["encoder_layers"]
    return conv_ln + 4 * spec[-1][0] * T + (4 * T * D + 6 * M * D) + 2 * Ly * M * D * 6 + 6 * T * D


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="large_64x20s", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-incremental", action="store_true", help="skip the incremental-mode latency probe")
    ap.add_argument("--no-extra-points", action="store_true", help="N > 1: skip the 30 s weak / global-128 strong points")
    ap.add_argument("--kernel-detail", action="store_true", help="per-kernel (name, shape) device ms in the JSON line")
    ap.add_argument("--cpu-sample-batch", type=int, default=4,
                    help="utterances per step of the CPU reference arm (a bounded sample of the workload's batch)")
    ap.add_argument("--step-blocks", type=int, default=1, help="streaming workloads: blocks of 16 frames per decision step")
    ap.add_argument("--step-impl", type=int, default=0, help="streaming workloads: 0 = automatic (cluster step kernel for one stream of the large model, else the operator chain), 1 = operator chain, 2 = first persistent step kernel")
    a = ap.parse_args()
    if a.workload.startswith("stream_"):
        kind, B, seconds = WORKLOADS[a.workload]
        if int(os.environ.get("RANK", "0")) == 0:
            stream_bench(a, kind, B, seconds)
        return
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    kind, B, seconds = WORKLOADS[a.workload]
    cfg = model_cfg(kind)
    L = seconds * SR
    metric = f"audio-sec encoded/sec (wav2vec-S {kind}, {a.dtype})"
    config = {"workload": f"wav2vec-S {kind} encoder forward, batch {B} x {seconds} s per GPU, {a.dtype}, "
                          f"main_context 16 / right_context 8, random-init weights",
              "name": a.workload, "batch_per_gpu": B, "utterance_s": seconds,
              "l2_policy": "inputs_larger_than_l2 (activations and weights exceed the 126 MB L2)"}

    if a.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(kind, min(a.cpu_sample_batch, B), seconds, a.steps, a.warmup)
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": "audio-s/s",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import wav2vec_s_b200 as W
    from wav2vec_s_b200 import cabi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float32

    torch.manual_seed(0)
    model = W.Wav2VecSModel(cfg)
    # random-init weights of the reference's scales, non-trivial biases / affine parameters
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for n_, p in model.named_parameters():
            if n_.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            elif p.dim() == 1 and n_.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif "feature_extractor" not in n_ and p.dim() == 2:
                p.copy_(torch.randn(p.shape, generator=g) * 0.04)
    model = model.to(dev, dtype).eval()

    gw = torch.Generator().manual_seed(1234 + rank)
    wav_host = torch.randn(B, L, generator=gw)
    wav_host = (wav_host - wav_host.mean(1, keepdim=True)) / wav_host.std(1, keepdim=True)
    wav_host = wav_host.pin_memory()
    wav_dev = wav_host.to(dev, non_blocking=True).to(dtype)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return model.extract_features(wav_dev, None)[0]

    # End to end through the public API with HOST buffers: every step copies its waveforms from pinned host
    # memory and reads its result back to pinned host memory.  The copies run on their own streams with two
    # buffers each, so the H2D of step i+1 and the D2H of step i-1 overlap the kernels of step i (what a user
    # feeding the encoder from a data loader does); all of it sits inside the timed region.
    main_stream = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    x_dev = [torch.empty_like(wav_host, device=dev) for _ in range(2)]
    out_host = [None, None]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]      # x_dev[k] consumed by the encoder
    ev_out = [torch.cuda.Event() for _ in range(2)]       # out_host[k] written
    e2e_i = [0]

    def step_e2e():
        i = e2e_i[0]; k = i & 1
        e2e_i[0] += 1
        with torch.cuda.stream(s_in):
            if i >= 2:
                s_in.wait_event(ev_free[k])
            x_dev[k].copy_(wav_host, non_blocking=True)
            ev_in[k].record(s_in)
        main_stream.wait_event(ev_in[k])
        y = model.extract_features(x_dev[k], None)[0]
        ev_free[k].record(main_stream)
        if out_host[k] is None:
            out_host[k] = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
        s_out.wait_stream(main_stream)
        with torch.cuda.stream(s_out):
            out_host[k].copy_(y, non_blocking=True)
            ev_out[k].record(s_out)
        y.record_stream(s_out)
        return y

    def drain_e2e():
        main_stream.wait_stream(s_out)
        main_stream.wait_stream(s_in)

    for _ in range(a.warmup):
        y = step_device()
    barrier()
    assert torch.isfinite(y.float()).all(), "non-finite encoder output"

    # ---- timed region: device-resident throughput -------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    cabi.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        step_device()
    e1.record()
    barrier()
    launches = cabi.launch_count(reset=True)
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region: end to end through the public API (pinned host in, host out) ---------
    for _ in range(2):
        step_e2e()
    drain_e2e()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(a.steps):
        step_e2e()
    drain_e2e()          # the last read-back is part of the timed region
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    # ---- N > 1: the one collective of the data plane -- gather every rank's outputs (NCCL all-gather over NVLink) ----
    gather = None
    if world > 1:
        from wav2vec_s_b200 import sharding
        yb = step_device()
        nfr = torch.full((B,), yb.size(1), dtype=torch.int64, device=dev)
        for _ in range(2):
            sharding.gather_outputs(yb, nfr)
        # the collective alone, into a preallocated buffer (NVLink all-gather bandwidth) ...
        gbuf = yb.new_empty((world * yb.size(0),) + tuple(yb.shape[1:]))
        for _ in range(2):
            dist.all_gather_into_tensor(gbuf, yb)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        g0.record()
        for _ in range(reps):
            dist.all_gather_into_tensor(gbuf, yb)
        g1.record()
        barrier()
        gms = g0.elapsed_time(g1) / reps
        del gbuf
        # ... and the public call (shape / count exchange, two host reads, per-utterance views)
        outs, counts = sharding.gather_outputs(yb, nfr)
        assert len(outs) == B * world and int(counts.sum()) == B * world * yb.size(1)
        # step + gather back to back (what a consumer that needs every rank's output on every rank pays)
        g2, g3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g2.record()
        for _ in range(a.steps):
            sharding.gather_outputs(step_device(), nfr)
        g3.record()
        barrier()
        tg = torch.tensor([gms, g2.elapsed_time(g3)], device=dev, dtype=torch.float64)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        rx = yb.numel() * yb.element_size() * (world - 1)
        gather = {"what": "all-gather of every rank's [B, T, D] outputs (NCCL all_gather_into_tensor; `ms` = the collective "
                          "alone, `ms_per_step_with_gather` = encoder step + sharding.gather_outputs back to back)",
                  "ms": float(tg[0]), "bytes_received_per_rank": rx, "gbps_per_rank": rx / (float(tg[0]) / 1e3) / 1e9,
                  "ms_per_step_with_gather": float(tg[1]) / a.steps,
                  "value_with_gather": B * seconds * world / (float(tg[1]) / a.steps / 1e3)}
        del outs, yb

    # ---- N > 1: two more points of SURVEY.md 8(d) on the same ranks: 30 s utterances (weak) and a fixed global batch
    #      of 128 x 20 s (strong) ----
    extra_points = None
    if world > 1 and a.workload == "large_64x20s" and not a.no_extra_points:
        extra_points = {}
        for name, (bb, secs) in {"large_64x30s_weak": (64, 30), "large_global128x20s_strong": (max(1, 128 // world), 20)}.items():
            gx = torch.Generator().manual_seed(99 + rank)
            xw = torch.randn(bb, secs * SR, generator=gx).to(dev).to(dtype)
            for _ in range(2):
                model.extract_features(xw, None)
            barrier()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record()
            for _ in range(5):
                model.extract_features(xw, None)
            x1.record()
            barrier()
            tx = torch.tensor([x0.elapsed_time(x1)], device=dev, dtype=torch.float64)
            dist.all_reduce(tx, op=dist.ReduceOp.MAX)
            extra_points[name] = {"batch_per_gpu": bb, "utterance_s": secs, "ms_per_step": float(tx[0]) / 5,
                                  "value": bb * secs * world / (float(tx[0]) / 5 / 1e3), "unit": "audio-s/s"}
            del xw
        model._ws = None

    # ---- per-kernel-class device time of one step (CUDA events on the launching stream) -----
    prof = None
    if hasattr(cabi.lib(), "w2vs_prof_enable"):
        prof = cabi.profile_step(step_device, reps=2, detail=a.kernel_detail)
    barrier()

    if rank == 0:
        audio_s = B * seconds * world
        ms_step = ms / a.steps
        fl = flops_per_utt(cfg, L)
        pk, pk_kind = peaks()
        line = {
            "metric": metric, "value": audio_s / (ms_step / 1e3), "unit": "audio-s/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic", "config": config,
            "clocks": clocks,
            "e2e": {"value": audio_s / (ms_e2e / a.steps / 1e3), "unit": "audio-s/s",
                    "h2d_bytes_per_step": wav_host.numel() * wav_host.element_size(),
                    "d2h_bytes_per_step": out_host[0].numel() * out_host[0].element_size(),
                    "pipeline": "double-buffered copy streams (H2D of step i+1 and D2H of step i-1 overlap step i)"},
            "gpu_launches": int(launches),
            "step_tflops": fl["total"] * B / (ms_step / 1e3) / 1e12,
            "step_frac_of_bf16_sustained": fl["total"] * B / (ms_step / 1e3) / 1e12 / pk["bf16_tflops_sustained"],
        }
        if prof is not None and prof.get("gemm", {}).get("count"):
            gm = prof["gemm"]
            ach = fl["gemm"] * B / (gm["ms"] / 1e3) / 1e12
            # DRAM traffic per class and step: dram__bytes_read.sum + dram__bytes_write.sum of EVERY launch of one step,
            # one ncu pass of this command (tools/ncu_traffic.py; a profiler cannot run inside the timed process)
            traffic, traffic_note, tclass = None, None, {}
            for tname in ("r02_traffic_" + a.workload + ".json",):
                tpath = os.path.join(ROOT, "profiles", tname)
                if os.path.isfile(tpath):
                    tj = json.load(open(tpath))
                    tclass = tj.get("bytes_per_step", {})
                    traffic = tclass.get("gemm")
                    wbytes = cfg["encoder_layers"] * (4 * cfg["encoder_embed_dim"] ** 2 + 2 * cfg["encoder_embed_dim"] * cfg["encoder_ffn_embed_dim"]) * 2
                    traffic_note = (f"{tj.get('source', 'ncu')}; launches per step covered: {tj.get('launches_per_step')}; "
                                    f"algorithmic bytes of the GEMM class (activations once + weights once per launch): "
                                    f"{int(gemm_bytes_per_utt(cfg, L) * B + wbytes)}")
            line["roofline"] = {"bound": "tensor", "kernel": "gemm_tc2_kernel (all launches of one step)",
                                "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                "frac": ach / pk["bf16_tflops_sustained"], "traffic": traffic,
                                "traffic_note": traffic_note,
                                "peak_source": pk_kind + " (sustained: timed inside a long step)",
                                "launches_per_step": gm["count"], "ms_per_step": gm["ms"]}
            line["kernel_ms_per_step"] = {k: round(v["ms"], 3) for k, v in prof.items() if "[" not in k and "_kernel" not in k}
            # every kernel class against the roofline that bounds it (tensor pipe: sustained bf16 peak; HBM: copy peak)
            rl = [dict(line["roofline"], **{"class": "gemm"})]
            if prof.get("attention", {}).get("ms"):
                t_at = prof["attention"]["ms"]
                rl.append({"class": "attention", "kernel": "attn_tc_kernel (block-mask-visible pairs only)", "bound": "tensor",
                           "achieved": fl["attn"] * B / (t_at / 1e3) / 1e12, "peak": pk["bf16_tflops_sustained"],
                           "unit": "TFLOP/s", "frac": fl["attn"] * B / (t_at / 1e3) / 1e12 / pk["bf16_tflops_sustained"],
                           "ms_per_step": t_at, "traffic": tclass.get("attention")})
            if prof.get("rows", {}).get("ms"):
                rb = rows_bytes_per_utt(cfg, L) * B
                rl.append({"class": "rows", "kernel": "LayerNorm / LN+GELU / embed / finalize passes", "bound": "hbm",
                           "achieved": rb / (prof["rows"]["ms"] / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                           "frac": rb / (prof["rows"]["ms"] / 1e3) / 1e9 / pk["hbm_gbs"], "ms_per_step": prof["rows"]["ms"],
                           "traffic": tclass.get("rows"), "algorithmic_bytes": rb})
            if prof.get("conv0", {}).get("ms"):
                cb = (wav_dev.element_size() * L + 2 * conv_spec(cfg)[0][0] * ((L - 10) // 5 + 1)) * B
                rl.append({"class": "conv0", "kernel": "conv0_kernel (Conv1d 1->512 k10 s5 + LayerNorm + GELU)", "bound": "hbm",
                           "achieved": cb / (prof["conv0"]["ms"] / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                           "frac": cb / (prof["conv0"]["ms"] / 1e3) / 1e9 / pk["hbm_gbs"], "ms_per_step": prof["conv0"]["ms"],
                           "traffic": tclass.get("conv0"), "algorithmic_bytes": cb})
            line["rooflines"] = rl
            if a.kernel_detail:
                line["kernel_detail"] = {k: [round(v["ms"], 3), v["count"]] for k, v in sorted(prof.items())
                                         if "[" in k or "_kernel" in k}
        else:
            line["roofline"] = None     # fp32 parity mode: CUDA-core kernels, no tensor-pipe roofline claimed
            if prof is not None:
                line["kernel_ms_per_step"] = {k: round(v["ms"], 3) for k, v in prof.items() if "[" not in k and "_kernel" not in k}
        if kind == "large" and world == 1 and not a.no_incremental:
            # the other half of BASELINE.json's metric: p50 latency of one decision step (16 new frames) of the
            # incremental path, one stream, same model; measured after the timed regions above
            try:
                line["incremental"] = incremental_probe(model, cfg, dev, B=1)
                line["incremental_b16"] = incremental_probe(model, cfg, dev, B=16)
                if line.get("rooflines") is not None:
                    for key in ("incremental", "incremental_b16"):
                        line["rooflines"].append(dict(line[key]["roofline"], **{"class": key, "ms_per_step": line[key]["p50_ms"]}))
            except Exception as ex:   # never lose the headline line to the secondary measurement
                line["incremental"] = {"error": f"{type(ex).__name__}: {ex}"}
        if gather is not None:
            line["gather"] = gather
        if extra_points is not None:
            line["extra_points"] = extra_points
        if not a.no_cpu_baseline and world == 1:      # reported on rank 0 at N=1 only
            r = cpu_reference_run(kind, min(a.cpu_sample_batch, B), seconds, 3, 1, bf16_too=True)   # ~25 s of CPU work at cfg3
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "bf16_value") if k in r}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
