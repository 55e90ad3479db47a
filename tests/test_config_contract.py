"""Config contract: every JSON config of the reference parses into the argument dataclasses (dmi/utils/args.py), with the
same splitting rule as HfArgumentParser.parse_json_file, and the derived rules of args_post_init hold."""
import dataclasses
import json
import os
import tempfile

import pytest

from dmi_b200.utils import args as A


@pytest.fixture(scope="module")
def configs(golden_dir):
    return json.load(open(os.path.join(golden_dir, "config_contract.json")))


def classes_for(rel):
    top = rel.split(os.sep)[0]
    return {"hypernet": A.HYPERNET_ARG_CLASSES, "projector": A.PROJECTOR_ARG_CLASSES, "lora": A.LORA_ARG_CLASSES}[top]


def test_all_reference_configs_parse(configs):
    assert len(configs) == 74
    counts = {"hypernet": 0, "projector": 0, "lora": 0}
    for rel, data in configs.items():
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
            json.dump(data, f)
        try:
            parsed = A.parse_json_file(f.name, classes_for(rel))
        finally:
            os.unlink(f.name)
        flat = {}
        for obj in parsed:
            flat.update(dataclasses.asdict(obj))
        for k, v in data.items():
            got = flat[k]
            assert (list(got) if isinstance(got, tuple) else got) == v, (rel, k)
        counts[rel.split(os.sep)[0]] += 1
    assert counts == {"hypernet": 19, "projector": 37, "lora": 18}


def test_matches_hf_argument_parser(configs):
    """the dataclasses are accepted by transformers.HfArgumentParser itself (what train_hypernet.py:653-661 uses)"""
    from transformers import HfArgumentParser
    rel = "hypernet/v4:llama1b_inst_all.json"
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(configs[rel], f)
    try:
        hf = HfArgumentParser(A.HYPERNET_ARG_CLASSES).parse_json_file(json_file=f.name)
        ours = A.parse_json_file(f.name, A.HYPERNET_ARG_CLASSES)
    finally:
        os.unlink(f.name)
    for a, b in zip(hf, ours):
        assert dataclasses.asdict(a) == dataclasses.asdict(b)


def test_defaults_and_field_order_match_the_reference_contract():
    t = A.TrainArgs(output_dir="x")
    assert (t.mode, t.device, t.train_batch_size, t.subset_batch_size, t.adam_beta2, t.warmup_steps, t.scheduler) == ("train", "mps", 128, 128, 0.999, 500, "cosine_warmup")
    assert t.seeds == (55625, 66848, 92900, 5225, 71753) and t.gradient_accumulation_steps == 1 and t.augment_emb_space is False
    h = A.HypnetArgs()
    assert (h.hn_arch, h.hn_n_heads, h.hn_hypnet_dim, h.hn_rank, h.hn_alpha, h.hn_predict_bias, h.hn_use_pos_encs) == ("transformer", 1, 768, 32, 32, True, False)
    p = A.ProjectorArgs()
    assert (p.proj_arch, p.proj_act, p.proj_n_layers, p.proj_dropout, p.proj_prune) == ("mlp", "quick_gelu", 2, 0.1, None)
    assert [f.name for f in dataclasses.fields(A.FewshotArgs)][:3] == ["finetune_generated_projector", "fewshot_learning_rate", "fewshot_weight_decay"]
    with pytest.raises(TypeError):
        A.TrainArgs()                     # output_dir is required


def test_setup_args_and_post_init_rules():
    class Box:
        pass
    b = Box()
    A.setup_args(b, "hn_", A.HypnetArgs(hn_rank=8))
    assert b.rank == 8 and b.arch == "transformer" and not hasattr(b, "hn_rank")
    # train_hypernet.py:465-472
    tr, me, pr, hn = A.TrainArgs(output_dir="x", finetune_mm_dim=768), A.MEncArgs(["e"], [True], mm_dim=512), A.ProjectorArgs(), A.HypnetArgs()
    A.hypernet_args_post_init(tr, me, pr, hn)
    assert hn.hn_n_proj_layers == 2 and pr.proj_prune == 512 and me.mm_dim == 512
    tr, me, pr = A.TrainArgs(output_dir="x", finetune_mm_dim=768), A.MEncArgs(["e"], [True], mm_dim=1024), A.ProjectorArgs()
    A.hypernet_args_post_init(tr, me, pr, hn)
    assert pr.proj_prune is None and tr.n_components == 768 and me.mm_dim == 768
    with pytest.raises(ValueError):
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
            json.dump({"output_dir": "x", "not_a_key": 1}, f)
        A.parse_json_file(f.name, (A.TrainArgs,))
