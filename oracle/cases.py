"""TEST INFRASTRUCTURE ONLY -- the parity cases shared by the golden generator and the tests.

Each case is a pure function of its entry here: weights from ``synth.make_state_dict(cfg, wseed)``,
waveform from ``synth.make_waveform(B, L, xseed)``, optional ragged lengths.  Edge cases follow
SURVEY.md section 8(c): T<16, odd T (pad_length=1), ragged lengths, right_context=0, rc window
overrunning T', main/rc in the sampling range, extractor default vs layer_norm, pos sin vs conv,
layer_norm_first both, rain is_infer +/- finished, streaming prefixes 16k+8.
"""
from .w2vs_oracle import default_cfg, base_cfg, large_cfg

TINY_CONV = "[(64,10,5)] + [(64,3,2)]*4 + [(64,2,2)]*2"


def tiny(**over):
    cfg = default_cfg(extractor_mode="layer_norm", encoder_layers=3, encoder_embed_dim=128,
                      encoder_ffn_embed_dim=256, encoder_attention_heads=2,
                      conv_feature_layers=TINY_CONV)
    cfg.update(over)
    return cfg


CASES = {
    # name: dict(cfg, B, L, ragged, api, kwargs)
    "tiny_postln_ln1": dict(cfg=tiny(encoder_layers=12), B=2, L=8000),          # base-style: LN on conv0 only
    "tiny_preln_bias_ragged": dict(cfg=tiny(layer_norm_first=True, conv_bias=True), B=3, L=12000,
                                   ragged=True),                                # T=37 odd, ragged
    "tiny_groupnorm_posconv": dict(cfg=tiny(extractor_mode="default", pos_type="conv",
                                            encoder_layers=2), B=2, L=8000),
    "tiny_groupnorm_posconv_ragged": dict(cfg=tiny(extractor_mode="default", pos_type="conv",
                                                   encoder_layers=2, layer_norm_first=True),
                                          B=2, L=9000, ragged=True),
    "tiny_rc0": dict(cfg=tiny(right_context=0), B=2, L=8000),
    "tiny_ctx8_4": dict(cfg=tiny(main_context=8, right_context=4, layer_norm_first=True), B=2, L=10000,
                        ragged=True),
    "tiny_ctx32_16": dict(cfg=tiny(main_context=32, right_context=16), B=1, L=24000),
    # block sizes that are not powers of two (context_type="sampling" draws them, wav2vec_S.py:392-395)
    "tiny_ctx20_10": dict(cfg=tiny(main_context=20, right_context=10, layer_norm_first=True), B=2, L=16000,
                          ragged=True),
    "tiny_ctx12_6": dict(cfg=tiny(main_context=12, right_context=6), B=2, L=9000),
    "tiny_short_T12": dict(cfg=tiny(), B=2, L=4000),                             # T=12 < main
    "tiny_T1": dict(cfg=tiny(layer_norm_first=True), B=1, L=400),                # T=1
    "tiny_T17": dict(cfg=tiny(), B=1, L=5760),                                   # T=17 odd, rc overrun
    "tiny_T25_ragged": dict(cfg=tiny(layer_norm_first=True), B=4, L=8320, ragged=True),
    "tiny_rain_full": dict(cfg=tiny(layer_norm_first=True), B=2, L=8000, api="rain", ragged=True),
    "tiny_rain_infer": dict(cfg=tiny(layer_norm_first=True), B=1, L=12880, api="rain",
                            kwargs=dict(is_infer=True, finished=False)),         # T=40=16*2+8
    "tiny_rain_infer_finished": dict(cfg=tiny(), B=1, L=12880, api="rain",
                                     kwargs=dict(is_infer=True, finished=True)),
    "tiny_stream_preln": dict(cfg=tiny(layer_norm_first=True, conv_bias=True), B=1, L=32000, api="stream"),
    "tiny_stream_postln": dict(cfg=tiny(encoder_layers=12), B=1, L=25000, api="stream"),
    "base_1s": dict(cfg=base_cfg(), B=2, L=16000, ragged=True),
    "large_1s": dict(cfg=large_cfg(), B=1, L=20000),
    # BASELINE.json shapes.  `compact=k`: the golden file keeps the final output only, every k-th frame (k = 3 is
    # coprime with the 16-frame blocks, so every position inside a block and every block is sampled) -- the full
    # tensors would put tens of MB into the repository.
    "cfg1_base_10s": dict(cfg=base_cfg(), B=1, L=160000, compact=1),            # configs[0] exactly: T=499, M=748
    "large_20s": dict(cfg=large_cfg(), B=1, L=320000, compact=3),               # configs[2] utterance: T=999, M=1496
}
# golden cases small enough for the stage-by-stage tests (the compact ones have their own tests)
SMALL = [n for n, c in CASES.items() if not c.get("compact")]

WSEED = 7
XSEED = 1234
LSEED = 4321
