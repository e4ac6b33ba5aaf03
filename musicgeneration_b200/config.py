"""Run-time constants read by the host mirror (the role MT/config.py:9-55 plays upstream).

Only the integers the hot path consumes are kept.  Upstream derives the vocabulary from its
MIDI-like tokenizer (88 note-on + 88 note-off + 32 velocity + 100 time-shift = 308 events,
MT/sequence.py:12,20-21,195-203; pad id = 308; vocabulary = 309) -- stated here as numbers so
nothing depends on pretty_midi.  ``MusicTransformer.forward`` looks up ``config.pad_token`` at
call time (as MT/network.py:37 does), and ``generate`` looks up ``config.threshold_len``
(MT/network.py:53), so callers may overwrite either before a call.
"""
import torch

# --- vocabulary -------------------------------------------------------------------------
_NOTE_ON, _NOTE_OFF, _VELOCITY, _TIME_SHIFT = 88, 88, 32, 100
event_dim = _NOTE_ON + _NOTE_OFF + _VELOCITY + _TIME_SHIFT
pad_token = event_dim
vocab_size = event_dim + 1

# --- model hyper-parameters (ctor keywords of MusicTransformer) ---------------------------
embedding_dim, num_layers, max_seq, dropout = 256, 6, 2048, 0.2
model = dict(vocab_size=vocab_size, embedding_dim=embedding_dim, max_seq=max_seq,
             num_layer=num_layers, dropout=dropout)

# --- optimisation loop ("next" row 1 of SURVEY section 8f) --------------------------------
batch_size, accum_grad, label_smooth, l_r, epochs = 6, 12, 0.1, 1e-4, 50000
warmup_steps, adam_betas, adam_eps = 4000, (0.9, 0.98), 1e-9

# --- sampling -----------------------------------------------------------------------------
length, threshold_len = 2000, 500

device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
debug = False
