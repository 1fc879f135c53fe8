"""Side-effect-free reader for the reference's experiment .cfg files.

Produces the same attribute bag as the reference's `data.read_config` (data.py:19-130) -- same
names, types and derived fields -- but never creates folders or shells out, so benchmarks and
tests can build models where the reference tree (and its datasets) are absent.  When the
reference's own `main.py` drives training it keeps using its own `data.read_config`; both
produce objects `models.Model` / `models.PretrainedModel` accept.
"""
import configparser


class Config:
    def __init__(self):
        self.use_sincnet = True


def _ints(s):
    return [int(v) for v in s.split(",")]


def _floats(s):
    return [float(v) for v in s.split(",")]


def _strs(s):
    return [v for v in s.split(",")]


def _flag(s):
    return s == "True"


_FIELDS = [
    # (section, key, attribute, converter)
    ("experiment", "seed", "seed", int), ("experiment", "folder", "folder", str),
    ("phoneme_module", "use_sincnet", "use_sincnet", _flag), ("phoneme_module", "fs", "fs", int),
    ("phoneme_module", "cnn_N_filt", "cnn_N_filt", _ints), ("phoneme_module", "cnn_len_filt", "cnn_len_filt", _ints),
    ("phoneme_module", "cnn_stride", "cnn_stride", _ints), ("phoneme_module", "cnn_max_pool_len", "cnn_max_pool_len", _ints),
    ("phoneme_module", "cnn_act", "cnn_act", _strs), ("phoneme_module", "cnn_drop", "cnn_drop", _floats),
    ("phoneme_module", "phone_rnn_num_hidden", "phone_rnn_num_hidden", _ints),
    ("phoneme_module", "phone_downsample_len", "phone_downsample_len", _ints),
    ("phoneme_module", "phone_downsample_type", "phone_downsample_type", _strs),
    ("phoneme_module", "phone_rnn_drop", "phone_rnn_drop", _floats),
    ("phoneme_module", "phone_rnn_bidirectional", "phone_rnn_bidirectional", _flag),
    ("word_module", "word_rnn_num_hidden", "word_rnn_num_hidden", _ints),
    ("word_module", "word_downsample_len", "word_downsample_len", _ints),
    ("word_module", "word_downsample_type", "word_downsample_type", _strs),
    ("word_module", "word_rnn_drop", "word_rnn_drop", _floats),
    ("word_module", "word_rnn_bidirectional", "word_rnn_bidirectional", _flag),
    ("word_module", "vocabulary_size", "vocabulary_size", int),
    ("intent_module", "intent_rnn_num_hidden", "intent_rnn_num_hidden", _ints),
    ("intent_module", "intent_downsample_len", "intent_downsample_len", _ints),
    ("intent_module", "intent_downsample_type", "intent_downsample_type", _strs),
    ("intent_module", "intent_rnn_drop", "intent_rnn_drop", _floats),
    ("intent_module", "intent_rnn_bidirectional", "intent_rnn_bidirectional", _flag),
    ("pretraining", "asr_path", "asr_path", str), ("pretraining", "pretraining_type", "pretraining_type", int),
    ("pretraining", "pretraining_lr", "pretraining_lr", float),
    ("pretraining", "pretraining_batch_size", "pretraining_batch_size", int),
    ("pretraining", "pretraining_num_epochs", "pretraining_num_epochs", int),
    ("pretraining", "pretraining_length_mean", "pretraining_length_mean", float),
    ("pretraining", "pretraining_length_var", "pretraining_length_var", float),
    ("training", "slu_path", "slu_path", str), ("training", "unfreezing_type", "unfreezing_type", int),
    ("training", "training_lr", "training_lr", float), ("training", "training_batch_size", "training_batch_size", int),
    ("training", "training_num_epochs", "training_num_epochs", int),
]
_OPTIONAL = [   # (section, key, attribute, converter, default)  -- data.py:95-119 try/except defaults
    ("training", "real_dataset_subset_percentage", "real_dataset_subset_percentage", float, 1.0),
    ("training", "synthetic_dataset_subset_percentage", "synthetic_dataset_subset_percentage", float, 1.0),
    ("training", "real_speaker_subset_percentage", "real_speaker_subset_percentage", float, 1.0),
    ("training", "synthetic_speaker_subset_percentage", "synthetic_speaker_subset_percentage", float, 1.0),
    ("training", "augment", "augment", _flag, False), ("training", "seq2seq", "seq2seq", _flag, False),
    ("training", "dataset_upsample_factor", "dataset_upsample_factor", int, 1),
]
_SEQ2SEQ = ["intent_encoder_dim", "num_intent_encoder_layers", "intent_decoder_dim", "num_intent_decoder_layers",
            "intent_decoder_key_dim", "intent_decoder_value_dim"]


def read_config(config_file):
    parser = configparser.ConfigParser()
    if not parser.read(config_file):
        raise FileNotFoundError(config_file)
    cfg = Config()
    for section, key, attr, conv in _FIELDS:
        setattr(cfg, attr, conv(parser.get(section, key)))
    for section, key, attr, conv, default in _OPTIONAL:
        setattr(cfg, attr, conv(parser.get(section, key)) if parser.has_option(section, key) else default)
    for opt in ("train_wording_path", "test_wording_path"):
        v = parser.get("training", opt) if parser.has_option("training", opt) else "None"
        setattr(cfg, opt, None if v == "None" else v)
    try:   # all-or-nothing like the reference's try block (data.py:66-74)
        vals = {k: int(parser.get("intent_module", k)) for k in _SEQ2SEQ}
        for k, v in vals.items():
            setattr(cfg, k, v)
    except (configparser.NoOptionError, ValueError):
        pass
    n_pre = {0: 1 + len(cfg.word_rnn_num_hidden) + len(cfg.phone_rnn_num_hidden) + len(cfg.cnn_N_filt),
             1: 1 + len(cfg.word_rnn_num_hidden), 2: 1, 3: 1}
    cfg.starting_unfreezing_index = n_pre[cfg.pretraining_type]                  # data.py:79-82
    cfg.phone_downsample_factor = 1
    for f in cfg.cnn_stride + cfg.cnn_max_pool_len + cfg.phone_downsample_len:
        cfg.phone_downsample_factor *= f
    cfg.word_downsample_factor = cfg.phone_downsample_factor
    for f in cfg.word_downsample_len:
        cfg.word_downsample_factor *= f
    return cfg


def fsc_intent_table():
    """values_per_slot / Sy_intent shape of Fluent Speech Commands (6 actions, 14 objects, 4 locations);
    the CSVs are not redistributable, so labels are indices (data.py:191-200 builds the real table)."""
    names = ("action", "object", "location")
    sizes = (6, 14, 4)
    return {n: {"%s_%d" % (n, i): i for i in range(s)} for n, s in zip(names, sizes)}, list(sizes)
