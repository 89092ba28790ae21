"""AstarPa2Params of every oracle configuration (oracle/oracle_capi.cpp preset_params), as the GPU library takes them.

0 / 1 are the presets; 2..9 mirror the reference's own test matrix (astarpa2/src/tests.rs:19-119, minus the SH heuristic
and the ignored local_doubling test); 10..15 are further points of the parameter space."""


def params_for(apa, preset):
    P = apa.AstarPa2Params
    simple, full = P.simple(), P.full()
    # BlockParams::default() (astarpa2/src/blocks.rs:59-73): sparse, incremental_doubling, no dt_trace, max_g 40, fr_drop 20
    dflt = dict(sparse=1, incremental_doubling=1, dt_trace=0, max_g=40, fr_drop=20)
    band_gap = dict(doubling="band_doubling", doubling_start="gap", factor=2.0)  # DoublingType::band_doubling(), band.rs:48-53
    table = {
        0: simple,
        1: full,
        2: full.replace(k=15, p=0, **band_gap, **dflt),                                     # nw_prune
        3: full.replace(k=15, p=0, **band_gap, **dict(dflt, dt_trace=1)),                   # dt_trace
        4: simple.replace(block_width=64, prune=1, **band_gap, **dflt),                     # band_doubling_edlib
        5: simple.replace(block_width=64, prune=1, **band_gap, **dict(dflt, dt_trace=1)),   # incremental_doubling
        6: simple.replace(heuristic="none", block_width=64, prune=1, **band_gap, **dflt),   # band_doubling_dijkstra
        7: simple.replace(domain="gap_gap", block_width=64, prune=1, **band_gap, **dflt),   # band_doubling_gapgap
        8: simple.replace(domain="gap_gap", block_width=256, prune=1, **band_gap, **dict(dflt, dt_trace=1)),  # dt_trace_gapgap
        9: simple.replace(domain="full", doubling="none", block_width=1, prune=1, **dflt),  # full
        10: simple.replace(domain="gap_start", block_width=64, **band_gap, **dict(dflt, dt_trace=1)),
        11: simple.replace(doubling="linear_search", doubling_start="gap", delta=48, block_width=32),
        12: full.replace(sparse_h=0, block_width=32, doubling_start="zero", factor=1.5, fr_drop=20),
        13: simple.replace(domain="full", doubling="none", block_width=256, sparse_h=0, prune=0, incremental_doubling=0,
                           dt_trace=0, max_g=40, fr_drop=20),
        14: simple.replace(heuristic="none", block_width=1, prune=1, **band_gap, **dict(dflt, dt_trace=1)),
        15: full.replace(k=8, p=3, prune=0, block_width=100, doubling="linear_search", delta=200),
    }
    return table[preset]


GENERAL_PRESETS = list(range(16))
