"""TEST INFRASTRUCTURE ONLY -- dumps the merged ``model`` / optimizer settings of every config under the reference's
configs/exp_local/ as JSON fixtures (tests/golden/configs/*.json).

    python -m oracle.make_config_fixtures        # needs /root/reference; writes tests/golden/configs/

mmcv's Config.fromfile is not installed here, so the `_base_` inheritance is resolved by the ~30-line merger below, which
follows mmcv 1.3.18 semantics (Config._file2dict / _merge_a_into_b: bases are loaded first and must not share top-level
keys, the child is merged into them dict by dict, a non-dict value -- including None -- replaces the base value, a dict
carrying `_delete_=True` replaces instead of merging).  tests/test_configs_cpu.py builds every dumped model through
clover_b200.registry (the drop-in boundary of SURVEY.md 8b) and, when /root/reference is mounted, re-merges the original
files and checks the fixtures are current.
"""
import json
import os
import sys

REF_CONFIGS = "/root/reference/configs"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "configs")
KEEP = ("model", "optimizer", "optimizer_config", "fp16", "lr_config", "total_epochs", "find_unused_parameters")


def _merge(child, base):
    out = dict(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get("_delete_", False):
            out[k] = _merge(v, out[k])
        elif isinstance(v, dict):
            out[k] = {kk: vv for kk, vv in v.items() if kk != "_delete_"}
        else:
            out[k] = v
    return out


def load_config(path):
    ns = {}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), ns)
    cfg = {k: v for k, v in ns.items() if not k.startswith("__") and not callable(v) and not isinstance(v, type(os))}
    bases = cfg.pop("_base_", [])
    bases = [bases] if isinstance(bases, str) else bases
    merged = {}
    for b in bases:
        bc = load_config(os.path.normpath(os.path.join(os.path.dirname(path), b)))
        dup = set(merged) & set(bc)
        if dup:
            raise KeyError(f"duplicate keys in the bases of {path}: {sorted(dup)}")
        merged.update(bc)
    return _merge(cfg, merged)


def dump_all(src=REF_CONFIGS, out=OUT):
    os.makedirs(out, exist_ok=True)
    names = []
    for fn in sorted(os.listdir(os.path.join(src, "exp_local"))):
        if not fn.endswith(".py"):
            continue
        cfg = load_config(os.path.join(src, "exp_local", fn))
        keep = {k: cfg[k] for k in KEEP if k in cfg}
        keep["_source"] = f"configs/exp_local/{fn}"
        with open(os.path.join(out, fn[:-3] + ".json"), "w") as f:
            json.dump(keep, f, indent=1, sort_keys=True)
        names.append(fn)
    return names


if __name__ == "__main__":
    print("wrote", dump_all(), "to", OUT)
    sys.exit(0)
