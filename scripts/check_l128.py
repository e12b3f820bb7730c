import io, contextlib, sys
sys.path.insert(0, '.')
import gru4rec
from gru4rec_b200.synth import make_sessions
df = make_sessions(n_items=3000, n_events=60000, seed=1)
g = gru4rec.GRU4Rec(layers=[128], loss='bpr-max', final_act='elu-0.5', batch_size=32, n_sample=2048, n_epochs=1, learning_rate=0.05, momentum=0.1)
g.fit(df, sample_store=2048 * 64)
print('fast windows (fast, fallback):', g._engine.fast_windows(), 'step_mode', g._engine.cfg.step_mode)
