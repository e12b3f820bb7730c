from collections import OrderedDict
gru4rec_params = OrderedDict([
('layers', [32]),
('loss', 'bpr-max'),
('final_act', 'elu-0.5'),
('hidden_act', 'tanh'),
('adapt', 'adagrad'),
('n_epochs', 2),
('batch_size', 16),
('dropout_p_embed', 0.0),
('dropout_p_hidden', 0.0),
('learning_rate', 0.1),
('momentum', 0.3),
('sample_alpha', 0.5),
('n_sample', 64),
('bpreg', 1.0),
('constrained_embedding', False)
])
