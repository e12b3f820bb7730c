"""ms per mini-batch of the BASELINE.json configurations other than the headline (synthetic shapes from SURVEY.md 8d);
these are parity-test shapes, measured here only to show that the generic path covers them."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from gru4rec_b200 import _lib
from gru4rec_b200.synth import make_session_arrays
import gru4rec as g4

CONFIGS = {
    'cfg1 synthetic XE B=32 GRU(100) I=1k': (1000, dict(layers=[100], loss='cross-entropy', final_act='softmax', batch_size=32, n_sample=2048)),
    'cfg2 RSC15 BPR-max B=32 GRU(100)': (37483, dict(layers=[100], loss='bpr-max', final_act='elu-0.5', batch_size=32, learning_rate=0.2, momentum=0.3, sample_alpha=0.0, n_sample=2048)),
    "cfg2' RSC15 XE shared (paramfiles/rsc15_xe_shared_100_best.py)": (37483, dict(layers=[100], loss='cross-entropy', final_act='softmax', constrained_embedding=True, batch_size=32,
                                  dropout_p_hidden=0.4, learning_rate=0.2, momentum=0.2, n_sample=2048, sample_alpha=0.5, bpreg=0.0, logq=1.0)),
    'cfg3 Rees46 XE shared B=240 GRU(512)': (172000, dict(layers=[512], loss='cross-entropy', final_act='softmax', constrained_embedding=True, batch_size=240,
                                  dropout_p_embed=0.45, learning_rate=0.065, momentum=0.0, n_sample=2048, sample_alpha=0.5, bpreg=0.0, logq=1.0)),
    'cfg4 RetailRocket-shaped BPR-max shared 3xGRU(100) B=80': (37000, dict(layers=[100, 100, 100], loss='bpr-max', final_act='elu-0.5', constrained_embedding=True, batch_size=80,
                                  dropout_p_embed=0.5, dropout_p_hidden=0.05, learning_rate=0.05, momentum=0.4, n_sample=2048, sample_alpha=0.4, bpreg=1.95)),
}
K = 300
for name, (I, mk) in CONFIGS.items():
    B = mk['batch_size']
    cfg = _lib.make_config(I, mk, sample_store=2048 * 1000, max_resident_steps=K + 8, step_mode=2)
    eng = _lib.Engine(cfg)
    gru = g4.GRU4Rec(**mk); gru.n_items = I
    for n, w in gru._init_host_weights().items():
        eng.set(n, w)
    items, offset, order, supports = make_session_arrays(I, max(int(3 * K * B * 1.7), 4 * I), seed=0)
    P = supports.astype(np.float64) ** mk.get('sample_alpha', 0.75); P = P.cumsum() / P.sum(); P[-1] = 1
    eng.set_sampling_cdf(P.astype(np.float32))
    if mk.get('logq', 0):
        eng.set_logq_support(np.maximum(supports, 1).astype(np.float32))
    eng.generate_samples()
    sched = _lib.Schedule(items, offset, order, B, 2048, mode=0)
    eng.upload_steps(sched, 0, K); eng.run_uploaded(K, False)
    eng.upload_steps(sched, K, K); c, ms = eng.run_uploaded(K, True)
    assert np.isfinite(c).all()
    print('%-70s %8.1f us/mini-batch  %9.0f mb/s  %10.0f events/s  (fast windows %s)' % (name, ms / K * 1000, K / ms * 1000, K * B / ms * 1000, eng.fast_windows()))
    eng.close()
