"""Config contract of the reference (``dmi/utils/args.py:9-119``): the seven argument dataclasses that
``train_hypernet.py`` / ``train_projector.py`` / ``train_lora.py`` fill from one JSON file with ``HfArgumentParser``.

Field names, order and defaults are the contract (every JSON under the reference's ``dmi/configs`` must parse), so they
are kept exactly; the classes are generated from a compact table instead of being written out."""
from __future__ import annotations

import copy
import dataclasses
import json
from typing import List, Optional, Tuple

_REQ = dataclasses.MISSING

_SPEC = {
    "TrainArgs": [
        ("output_dir", str, _REQ), ("mode", str, "train"), ("device", str, "mps"),
        ("resume_from_checkpoint", str, None), ("finetune_from_checkpoint", str, None), ("finetune_mm_dim", int, None),
        ("resume_from_checkpoint_reset_steps", bool, False), ("save_state", bool, True),
        ("train_batch_size", int, 128), ("subset_batch_size", int, 128), ("eval_batch_size", int, 128),
        ("learning_rate", float, 1e-4), ("max_grad_norm", float, 1.0), ("weight_decay", float, 0.0),
        ("adam_beta1", float, 0.9), ("adam_beta2", float, 0.999), ("adam_epsilon", float, 1e-8),
        ("epochs", int, None), ("dataset_size", str, None), ("epochs_l", List[int], None), ("dataset_size_l", List[str], None),
        ("warmup_steps", int, 500), ("scheduler", str, "cosine_warmup"), ("logging_steps", int, 50),
        ("save_steps", int, 5000), ("save_steps_l", List[int], None), ("eval_steps", int, 5000), ("eval_steps_l", List[int], None),
        ("generate_steps", int, 5000), ("generate_steps_l", List[int], None),
        ("eval_at_step_zero", bool, False), ("generate_at_step_zero", bool, False), ("seed", int, 42),
        ("seeds", Tuple[int], (55625, 66848, 92900, 5225, 71753)),
        ("gradient_accumulation_steps", int, 1), ("pad_to_multiple_of", int, 8), ("debug", bool, False),
        ("feed_txt_embs", bool, False), ("augment_emb_space", bool, False), ("subtract_mean", bool, False),
        ("n_components", int, None),
    ],
    "MEncArgs": [
        ("menc_names_or_paths", List[str], _REQ), ("load_extracted_features", List[bool], _REQ),
        ("fewshot_menc_names_or_paths", List[str], None), ("fewshot_load_extracted_features", List[bool], None),
        ("mm_dim", int, 768), ("mm_dtype", Optional[str], "float32"),
    ],
    "LMArgs": [("lm_name_or_path", str, _REQ), ("lm_dtype", Optional[str], "bfloat16")],
    "DatasetArgs": [("dataset_names_or_paths", List[str], _REQ), ("fewshot_dataset_names_or_paths", List[str], None)],
    "ProjectorArgs": [
        ("proj_name_or_path", str, None), ("proj_arch", str, "mlp"), ("proj_act", str, "quick_gelu"),
        ("proj_n_layers", int, 2), ("proj_dropout", float, 0.1), ("proj_prune", int, None),
    ],
    "HypnetArgs": [
        ("hn_name_or_path", str, "hypnet_1"), ("hn_arch", str, "transformer"), ("hn_n_layers", int, 1), ("hn_n_heads", int, 1),
        ("hn_hypnet_dim", int, 768), ("hn_rank", int, 32), ("hn_alpha", int, 32), ("hn_predict_bias", bool, True),
        ("hn_principled_init", bool, False), ("hn_n_proj_layers", int, None), ("hn_use_pos_encs", bool, False),
    ],
    "LoraArgs": [
        ("lora_name_or_path", str, "lora_1"), ("lora_rank", int, 32), ("lora_alpha", int, 32), ("lora_n_proj_layers", int, None),
    ],
    "FewshotArgs": [
        ("finetune_generated_projector", bool, _REQ), ("fewshot_learning_rate", float, 1e-4), ("fewshot_weight_decay", float, 5e-6),
        ("fewshot_dataset_sizes", List[str], None), ("fewshot_epochs", List[int], None), ("fewshot_n_adapters", str, "multiple"),
        ("fewshot_n_tokens", int, None),
    ],
}


def _field(default):
    if default is _REQ:
        return dataclasses.field()
    if isinstance(default, (list, tuple, dict)):
        return dataclasses.field(default_factory=lambda d=default: copy.deepcopy(d))
    return dataclasses.field(default=default)


def _build(name):
    cls = dataclasses.make_dataclass(name, [(n, t, _field(d)) for n, t, d in _SPEC[name]])
    cls.__module__ = __name__
    return cls


TrainArgs = _build("TrainArgs")
MEncArgs = _build("MEncArgs")
LMArgs = _build("LMArgs")
DatasetArgs = _build("DatasetArgs")
ProjectorArgs = _build("ProjectorArgs")
HypnetArgs = _build("HypnetArgs")
LoraArgs = _build("LoraArgs")
FewshotArgs = _build("FewshotArgs")

HYPERNET_ARG_CLASSES = (TrainArgs, MEncArgs, LMArgs, DatasetArgs, ProjectorArgs, HypnetArgs, FewshotArgs)   # train_hypernet.py:653
PROJECTOR_ARG_CLASSES = (TrainArgs, MEncArgs, LMArgs, DatasetArgs, ProjectorArgs)                             # train_projector.py:299
LORA_ARG_CLASSES = (TrainArgs, MEncArgs, LMArgs, DatasetArgs, ProjectorArgs, LoraArgs)                       # train_lora.py


def setup_args(obj, prefix: str, args) -> None:
    """copy every ``<prefix>foo`` attribute of ``args`` onto ``obj.foo`` (reference ``setup_args``, args.py:116-120)"""
    for key in dir(args):
        if key.startswith(prefix):
            setattr(obj, key[len(prefix):], getattr(args, key))


def parse_json_file(path: str, classes) -> tuple:
    """Same splitting rule as ``HfArgumentParser.parse_json_file``: each key goes to the dataclass that declares it;
    unknown keys are an error."""
    with open(path) as f:
        data = json.load(f)
    used, out = set(), []
    for cls in classes:
        names = {f.name for f in dataclasses.fields(cls)}
        out.append(cls(**{k: v for k, v in data.items() if k in names}))
        used |= names & set(data)
    extra = set(data) - used
    if extra:
        raise ValueError(f"Some keys are not used by the argument dataclasses: {sorted(extra)}")
    return tuple(out)


def hypernet_args_post_init(train_args, menc_args, proj_args, hn_args) -> None:
    """Derived rules of ``train_hypernet.py:465-472``: hn_n_proj_layers := proj_n_layers; pruning / n_components from
    finetune_mm_dim vs mm_dim."""
    hn_args.hn_n_proj_layers = proj_args.proj_n_layers
    ft = train_args.finetune_mm_dim
    if ft is not None:
        if menc_args.mm_dim < ft:
            proj_args.proj_prune = menc_args.mm_dim
        elif menc_args.mm_dim > ft:
            train_args.n_components = ft
            menc_args.mm_dim = ft
