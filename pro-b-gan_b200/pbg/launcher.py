"""Run the *unmodified* reference entry point on top of a ``modular_prot_b_gan`` module.

``pro_b_gan_infer.py`` does ``from modular_prot_b_gan import ModularGenerator,
ModularDiscriminator`` (:41) and then instantiates the undefined names
``Generator`` / ``Discriminator`` (:93-94), so as shipped it dies with NameError.
The drop-in seam is therefore:

  1. a module importable as ``modular_prot_b_gan`` (this package ships the
     CUDA-backed one next to this file's parent directory), and
  2. the two names injected into the script's globals before
     ``ProtBGANInference.__init__`` runs.

Usage (same CLI as the reference, :437-461):

    python -m pbg.launcher /path/to/pro_b_gan_infer.py --checkpoint_path ckpt.pt \
        --task score_triplets --input_triplets "[[0,1,2]]"
"""
from __future__ import annotations

import importlib
import importlib.util
import sys
from types import ModuleType


def load_reference_script(script_path: str, model_module: ModuleType | None = None,
                          module_name: str = "pro_b_gan_infer_ref") -> ModuleType:
    """Import ``script_path`` untouched with ``model_module`` visible as ``modular_prot_b_gan``.

    Returns the imported script module with ``Generator`` / ``Discriminator`` set.
    """
    if model_module is None:
        model_module = importlib.import_module("modular_prot_b_gan")
    saved = sys.modules.get("modular_prot_b_gan")
    sys.modules["modular_prot_b_gan"] = model_module
    try:
        spec = importlib.util.spec_from_file_location(module_name, script_path)
        if spec is None or spec.loader is None:
            raise FileNotFoundError(f"cannot import reference script: {script_path}")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        if saved is None:
            sys.modules.pop("modular_prot_b_gan", None)
        else:
            sys.modules["modular_prot_b_gan"] = saved
    ref.Generator = getattr(model_module, "Generator", model_module.ModularGenerator)
    ref.Discriminator = getattr(model_module, "Discriminator", model_module.ModularDiscriminator)
    return ref


def run_main(script_path: str, argv: list[str], model_module: ModuleType | None = None) -> None:
    ref = load_reference_script(script_path, model_module)
    saved_argv = sys.argv
    sys.argv = [script_path] + list(argv)
    try:
        ref.main()
    finally:
        sys.argv = saved_argv


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    run_main(sys.argv[1], sys.argv[2:])
