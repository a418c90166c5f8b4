"""Generate golden vectors by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py            # needs /root/reference (absent on the GPU box)

The reference modules `src/clip/eval/metrics.py` and `fusion.py` import with numpy only;
`src/retrieval.py` is imported with four stubs (its import logs in to the HF hub and needs
`mistralai` / `SPARQLWrapper`) and `RetrievalEngine.__new__` so that the
network-bound constructor never runs.  Outputs: `tests/golden/*.npz|json`, committed.
Inputs are produced by `knowledge_enhanced_multimodal_retrieval_b200.synth` and stored as
bf16 bit patterns so the fixtures do not depend on the generator staying unchanged.
"""
import contextlib
import io
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("KEMR_REFERENCE", "/root/reference")

from knowledge_enhanced_multimodal_retrieval_b200 import synth  # noqa: E402


def import_reference():
    sys.path.insert(0, REF)
    from src.clip.eval import metrics as rmetrics, fusion as rfusion
    import huggingface_hub
    huggingface_hub.login = lambda *a, **k: None
    huggingface_hub.hf_hub_download = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("offline"))
    for name, attrs in (("mistralai", ("Mistral",)), ("SPARQLWrapper", ("SPARQLWrapper", "JSON")),
                        ("dotenv", ())):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                for a in attrs:
                    setattr(m, a, object)
                if name == "dotenv":
                    m.load_dotenv = lambda *a, **k: None
                sys.modules[name] = m
    import src.retrieval as rretrieval
    return rmetrics, rfusion, rretrieval


class FakeTextDataset:
    """Stands in for TextOnlyDataset (evaluate_text_models.py:28-81): item i = the 5 text variants of artefact i."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return [f"t{i}v{v}" for v in range(5)]


class FakeTextModel:
    """Stands in for a SentenceTransformer: `.encode(texts, ...)` returns the stored (already normalised) vectors."""

    def __init__(self, table):
        self.table = table

    def to(self, device):
        return self

    def eval(self):
        return self

    def encode(self, texts, convert_to_tensor=True, device="cpu", show_progress_bar=False, normalize_embeddings=True):
        import torch
        return torch.from_numpy(np.stack([self.table[t] for t in texts]).astype(np.float32))


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def f64dict(d):
    return {k: float(v) for k, v in d.items()}


def main():
    rmetrics, rfusion, rretrieval = import_reference()
    out_json = {}

    # ------------------------------------------------------------------ metrics, small
    s = synth.make_retrieval_set(Q=96, M=160, D=64, seed=11, fused=True, lam=0.6, with_kg=True)
    q, img, tgt = s.query, s.image, s.target
    m = {}
    m["retrieval_metrics_T2I"] = f64dict(rmetrics.compute_retrieval_metrics(q, img, prefix="T2I"))
    m["retrieval_metrics_noprefix_k"] = f64dict(
        rmetrics.compute_retrieval_metrics(q, tgt, k_values=[1, 3, 7]))
    m["final_05_05"] = f64dict(quiet(rmetrics.compute_retrieval_metrics_final, q, tgt, img))
    m["final_01_09"] = f64dict(quiet(rmetrics.compute_retrieval_metrics_final, q, tgt, img,
                                     prefix="F", t2i_weight=0.1, t2t_weight=0.9))
    # the "all" variants need square inputs (I2T scores image i against target i)
    sq = synth.make_retrieval_set(Q=128, M=128, D=64, seed=12, fused=True, lam=0.5)
    m["all"] = f64dict(rmetrics.compute_all_retrieval_metrics(sq.query, sq.target, sq.image))
    m["all_T2I_T2T_mrr_only"] = f64dict(rmetrics.compute_all_retrieval_metrics(
        sq.query, sq.target, sq.image, tasks=["T2I", "T2T"], compute_recall=False))
    m["training"] = f64dict(rmetrics.compute_training_metrics(sq.query, sq.target, sq.image))
    # the three deprecated shims (metrics.py:285-352): first text variant against itself and against the images
    m["shim_multi_mode"] = f64dict(quiet(rmetrics.compute_metrics_multi_mode, sq.image, [sq.target, sq.query]))
    m["shim_single_4train"] = f64dict(quiet(rmetrics.compute_metrics_single_4train, sq.image, [sq.target]))
    m["shim_multi_4train"] = f64dict(quiet(rmetrics.compute_metrics_multi_4train, sq.image, [sq.query, sq.target]))
    sim = (q @ img.T).astype(np.float32)
    m["recall_at_k_matrix"] = f64dict(rmetrics.compute_recall_at_k(sim))
    m["mrr_matrix"] = f64dict(rmetrics.compute_mrr_and_mean_rank(sim))
    m["metrics_fusion_matrix"] = f64dict(rmetrics.compute_retrieval_metrics_fusion(sim, prefix="X"))
    m["evaluate_retrieval"] = f64dict(quiet(rfusion.evaluate_retrieval, sim))
    out_json["metrics_small"] = m

    # ------------------------------------------------------------------ dense KG fusion, small
    fus = {}
    args = (sim, s.kg_results, s.query_uuids, s.uuids)
    fus["weighted_default"] = rfusion.weighted_fusion(*args)
    fus["weighted_09_01"] = rfusion.weighted_fusion(*args, alpha=0.9, sparql_weight=1 - 0.9)
    fus["weighted_renorm"] = quiet(rfusion.weighted_fusion, *args, alpha=0.6, sparql_weight=0.6)
    fus["additive_default"] = rfusion.additive_bonus_fusion(*args)
    fus["additive_013"] = rfusion.additive_bonus_fusion(*args, delta=0.13)
    fus["adaptive_default"] = rfusion.adaptive_additive_fusion(*args)
    fus["adaptive_custom"] = rfusion.adaptive_additive_fusion(
        *args, delta=0.3, size_thresholds={2: 0.9, 10: 0.4, 25: 0.05})
    fus["dispatch_weighted"] = rfusion.fuse_clip_and_text2sparql(
        *args, fusion_strategy="weighted", fusion_params={"alpha": 0.4, "sparql_weight": 0.6})
    fus["dispatch_additive"] = rfusion.fuse_clip_and_text2sparql(*args, fusion_strategy="additive")
    fus["dispatch_adaptive"] = rfusion.fuse_clip_and_text2sparql(
        *args, fusion_strategy="adaptive", fusion_params={"delta": 0.25})
    fm = {}
    for name, mat in fus.items():
        assert mat.dtype == np.float32, (name, mat.dtype)
        fm[name] = f64dict(quiet(rfusion.evaluate_retrieval, mat))
    out_json["fusion_small_metrics"] = fm
    np.savez_compressed(
        os.path.join(HERE, "small_set.npz"),
        query_bits=synth.f32_to_bf16_bits(q), image_bits=synth.f32_to_bf16_bits(img),
        target_bits=synth.f32_to_bf16_bits(tgt), sim=sim,
        sq_query_bits=synth.f32_to_bf16_bits(sq.query), sq_image_bits=synth.f32_to_bf16_bits(sq.image),
        sq_target_bits=synth.f32_to_bf16_bits(sq.target),
        **{f"fusion_{k}": v for k, v in fus.items()})
    with open(os.path.join(HERE, "small_set_kg.json"), "w") as f:
        json.dump({"kg_results": s.kg_results, "query_uuids": s.query_uuids, "uuids": s.uuids}, f)

    # ------------------------------------------------------------------ metrics, mid size (seed-regenerated)
    mid = synth.make_retrieval_set(Q=1500, M=1500, D=128, seed=23, fused=True, lam=0.4, with_kg=True)
    mm = {"checksum": [float(mid.query.astype(np.float64).sum()), float(mid.image.astype(np.float64).sum()),
                       float(mid.target.astype(np.float64).sum())]}
    mm["all"] = f64dict(rmetrics.compute_all_retrieval_metrics(mid.query, mid.target, mid.image))
    for wi, wt in ((0.5, 0.5), (0.1, 0.9)):
        mm[f"final_{wi}_{wt}"] = f64dict(quiet(rmetrics.compute_retrieval_metrics_final,
                                               mid.query, mid.target, mid.image,
                                               t2i_weight=wi, t2t_weight=wt))
        fusedsim = (wi * (mid.query @ mid.image.T)) + (wt * (mid.query @ mid.target.T))
        for alpha in (0.9, 0.5, 0.1):            # evaluator.py:176-190 sweep (subset)
            fw = rfusion.fuse_clip_and_text2sparql(
                fusedsim, mid.kg_results, mid.query_uuids, mid.uuids, "weighted",
                {"alpha": alpha, "sparql_weight": 1 - alpha})
            mm[f"sweep_{wi}_{wt}_alpha{alpha}"] = f64dict(quiet(rfusion.evaluate_retrieval, fw))
    out_json["metrics_mid"] = mm
    # a set whose targets sit deep in the score bulk: fp32 near-ties next to the target exist,
    # so the reference's own numbers depend on its BLAS; kept to test that any deviation of
    # the canonical contract is confined to those audited rows
    nt = synth.make_retrieval_set(Q=1500, M=1500, D=128, seed=21, fused=True, lam=0.25)
    out_json["metrics_mid_nearties"] = {
        "final_0.5_0.5": f64dict(quiet(rmetrics.compute_retrieval_metrics_final, nt.query, nt.target, nt.image))}

    # ------------------------------------------------------------------ serving-engine list fusion
    eng = rretrieval.RetrievalEngine.__new__(rretrieval.RetrievalEngine)
    rng = np.random.default_rng(5)
    cases = []
    for n in (0, 1, 7, 40):
        clip = [{"uuid": f"a{j}", "score": float(np.float32(rng.uniform(-0.2, 0.9)))} for j in range(n)]
        clip.sort(key=lambda d: -d["score"])
        if n >= 7:                                  # exact and rounding-induced ties
            clip[3]["score"] = clip[2]["score"]
            clip[5]["score"] = clip[4]["score"] - 1e-6
        sparql = [f"a{j}" for j in rng.integers(0, max(n, 1), size=n // 3)] + ["zz-not-in-clip"]
        for alpha, beta in ((0.8, 0.2), (0.5, 0.5), (1.0, 0.0)):
            fused = eng._fuse_clip_sparql_linear(clip, sparql, alpha=alpha, beta=beta)
            cases.append({"clip": clip, "sparql": sparql, "alpha": alpha, "beta": beta, "out": fused})

    class FakeClip:
        def __init__(self, res):
            self.res = res

        def retrieval(self, query, alpha=0.5):
            return self.res

    class FakeT2S:
        def __init__(self, res):
            self.res = res

        def retrieval(self, query):
            return self.res

    eng.clip_retriever = FakeClip(cases[-1]["clip"])
    eng.t2s_retriever = FakeT2S(cases[-1]["sparql"])
    engine_calls = {
        "retrieve_text_default": eng.retrieve_text("q"),
        "retrieve_text_thr": eng.retrieve_text("q", alpha=0.6, beta=0.4, threshold=0.3),
        "noknowledge_default": eng.retrieve_text_noknowledge("q"),
        "noknowledge_thr": eng.retrieve_text_noknowledge("q", threshold=0.25),
    }
    out_json["engine"] = {"fuse_cases": cases, "calls": engine_calls,
                          "call_inputs": {"clip": cases[-1]["clip"], "sparql": cases[-1]["sparql"]}}

    # ------------------------------------------------------------------ learned fusion heads (fusion_model.py)
    # the UNMODIFIED torch modules of the reference, CPU fp32, eval mode, parameters set from a seeded generator;
    # scores -> the reference's own metrics.  Inputs: the square small set above (query i <-> candidate i).
    import torch
    from src.clip.model import fusion_model as rheads
    hr = np.random.default_rng(29)
    D = sq.query.shape[1]
    tq, ti, tt = (torch.from_numpy(np.ascontiguousarray(x)) for x in (sq.query, sq.image, sq.target))
    heads = {}

    def run_head(name, module, params):
        module.eval()
        with torch.no_grad():
            for key, val in params.items():
                obj, attr = module, key
                for part in key.split(".")[:-1]:
                    obj = getattr(obj, part) if not part.isdigit() else obj[int(part)]
                attr = key.split(".")[-1]
                getattr(obj, attr).copy_(torch.from_numpy(np.asarray(val, dtype=np.float32)).reshape(getattr(obj, attr).shape))
            scores = module(tq, ti, tt).numpy().astype(np.float32)
        heads[name] = {"params": {k: np.asarray(v, dtype=np.float32).reshape(-1).tolist() for k, v in params.items()},
                       "metrics": f64dict(rmetrics.compute_retrieval_metrics_fusion(scores)),
                       "top5": np.argsort(-scores, axis=1, kind="stable")[:, :5].tolist(),
                       "score_checksum": float(scores.astype(np.float64).sum()),
                       "score_row0": scores[0, :8].astype(np.float64).tolist()}
        return scores

    w = hr.normal(0, 0.6, D).astype(np.float32)
    run_head("simple_gated", rheads.SimpleGatedFusion(embed_dim=D), {"query_weight": w, "bias": [0.3]})
    run_head("simple_gated_with_bias", rheads.SimpleGatedFusionWithBias(embed_dim=D), {"query_weight": w * 0.5, "bias": -1.0})
    run_head("gated_mlp", rheads.GatedFusionHead(embed_dim=D),
             {"gate_net.0.weight": hr.normal(0, 0.2, (128, D)), "gate_net.0.bias": hr.normal(0, 0.1, 128),
              "gate_net.3.weight": hr.normal(0, 0.3, (1, 128)), "gate_net.3.bias": [0.05]})
    run_head("bilinear", rheads.BilinearFusionHead(embed_dim=D),
             {"W_image.weight": np.eye(D) + hr.normal(0, 0.02, (D, D)), "W_target.weight": np.eye(D) + hr.normal(0, 0.02, (D, D)),
              "alpha": 0.4})
    # LinearFusionHead takes the two similarity matrices (FusionModel.forward, fusion_model.py:318-322)
    lin = rheads.LinearFusionHead(hidden_dim=128)
    lin.eval()
    lp = {"fusion.0.weight": hr.normal(0, 0.8, (128, 2)), "fusion.0.bias": hr.normal(0, 0.2, 128),
          "fusion.3.weight": hr.normal(0, 0.3, (1, 128)), "fusion.3.bias": [0.02]}
    with torch.no_grad():
        for key, val in lp.items():
            mod = lin.fusion[int(key.split(".")[1])]
            getattr(mod, key.split(".")[2]).copy_(torch.from_numpy(np.asarray(val, dtype=np.float32)).reshape(getattr(mod, key.split(".")[2]).shape))
        lscores = lin(tq @ ti.T, tq @ tt.T).numpy().astype(np.float32)
    heads["linear"] = {"params": {k: np.asarray(v, dtype=np.float32).reshape(-1).tolist() for k, v in lp.items()},
                       "metrics": f64dict(rmetrics.compute_retrieval_metrics_fusion(lscores)),
                       "top5": np.argsort(-lscores, axis=1, kind="stable")[:, :5].tolist(),
                       "score_row0": lscores[0, :8].astype(np.float64).tolist(),
                       "score_checksum": float(lscores.astype(np.float64).sum())}
    # CrossAttentionFusionHead on a small sub-problem (it scores every PAIR with a 3-layer MLP)
    torch.manual_seed(31)
    ca = rheads.CrossAttentionFusionHead(embed_dim=D, num_heads=8, hidden_dim=256)
    ca.eval()
    with torch.no_grad():
        cscores = ca(tq[:48], ti[:48], tt[:48]).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "cross_attention_head.npz"),
                        **{k: v.detach().numpy() for k, v in ca.state_dict().items()}, scores=cscores)
    heads["cross_attention"] = {"metrics": f64dict(rmetrics.compute_retrieval_metrics_fusion(cscores)), "n": 48}
    with torch.no_grad():
        g = rheads.SimpleGatedFusion(embed_dim=D)
        g.query_weight.copy_(torch.from_numpy(w)); g.bias.fill_(0.3)
        gate = torch.sigmoid((tq * g.query_weight).sum(dim=1, keepdim=True) + g.bias).numpy().reshape(-1)
    heads["simple_gated"]["gate"] = gate.astype(np.float64).tolist()
    out_json["fusion_heads"] = heads

    # ------------------------------------------------------------------ contrastive losses (train/losses.py), forward values
    from src.clip.train import losses as rlosses
    with torch.no_grad():
        l1, m1 = rlosses.InfoNCELoss(temperature=0.07)(tq, tt)
        l2, m2 = rlosses.InfoNCELoss(temperature=0.5)(ti[:37], tt[:37])
        l3, m3 = rlosses.JointContrastiveLoss(temperature=0.07, t2i_weight=0.3, t2t_weight=0.9)(ti, tq, tt)
    out_json["losses"] = {"infonce_q_t_007": m1, "infonce_i_t_05_first37": m2, "joint_03_09": m3}

    # ------------------------------------------------------------------ grouped ground truth (baselines/evaluate_text_models.py)
    # the UNMODIFIED evaluate_text_model(), fed by a fake SentenceTransformer that looks embeddings up by text and a
    # fake 5-variant dataset; `sentence_transformers` itself is absent here and stubbed.
    if "sentence_transformers" not in sys.modules:
        st_mod = types.ModuleType("sentence_transformers")
        st_mod.SentenceTransformer = object
        sys.modules["sentence_transformers"] = st_mod
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_eval_text_models", os.path.join(REF, "baselines", "evaluate_text_models.py"))
    rtext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rtext)
    Ng, Dg = 48, 64
    gr = np.random.default_rng(41)
    base = synth.make_gallery(Ng, Dg, 43)
    variants = [synth.round_to_bf16(synth.l2_normalize(base * 0.7 + gr.normal(0, 1 / np.sqrt(Dg), (Ng, Dg)).astype(np.float32)))
                for _ in range(5)]
    table = {f"t{i}v{v}": variants[v][i] for i in range(Ng) for v in range(5)}
    grouped = {"variants_bits": [synth.f32_to_bf16_bits(v).tolist() for v in variants]}
    for mode in ("single", "multi"):
        res = quiet(rtext.evaluate_text_model, FakeTextModel(table), FakeTextDataset(Ng), batch_size=16, device="cpu",
                    mode=mode, seed=42)
        grouped[mode] = f64dict(res)
    out_json["grouped_text_models"] = grouped

    # ------------------------------------------------------------------ generic fp32 embeddings (NOT bf16-representable)
    # what a real CLIP encoder hands over: the engine rounds them to bf16 on upload, the reference scores them in fp32.
    # Recorded: the reference's metrics, its top-5 lists and a slice of its scores, so that the deviation caused by
    # the bf16 storage format is measured against the unmodified reference (tests/test_gpu_parity.py).
    gr2 = np.random.default_rng(77)
    Mg, Qg, Dg2 = 600, 200, 64
    gi = synth.l2_normalize(gr2.standard_normal((Mg, Dg2), dtype=np.float32))
    gt = synth.l2_normalize(gr2.standard_normal((Mg, Dg2), dtype=np.float32))
    gq = synth.l2_normalize(0.45 * (gi[:Qg] + gt[:Qg]) / 2 + gr2.standard_normal((Qg, Dg2), dtype=np.float32) / np.sqrt(Dg2))
    gsim = (0.5 * (gq @ gi.T)) + (0.5 * (gq @ gt.T))
    np.savez_compressed(os.path.join(HERE, "generic_fp32.npz"), query=gq, image=gi, target=gt)
    out_json["generic_fp32"] = {
        "final_0.5_0.5": f64dict(quiet(rmetrics.compute_retrieval_metrics_final, gq, gt, gi)),
        "T2I": f64dict(rmetrics.compute_retrieval_metrics(gq, gi)),
        "top5": np.argsort(-gsim, axis=1, kind="stable")[:, :5].tolist(),
        "top5_scores": np.take_along_axis(gsim, np.argsort(-gsim, axis=1, kind="stable")[:, :5], axis=1).astype(np.float64).tolist(),
    }
    # more queries than candidates (metrics.py:37 has no column i for i >= M): the reference's own answer
    tall = (q @ img[:40].T).astype(np.float32)
    out_json["tall_matrix"] = {"metrics": f64dict(rmetrics.compute_retrieval_metrics_fusion(tall)),
                               "embeddings": f64dict(rmetrics.compute_retrieval_metrics(q, img[:40]))}

    import numpy
    out_json["provenance"] = {"numpy": numpy.__version__, "reference": REF,
                              "note": "outputs of the unmodified reference functions"}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out_json, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
