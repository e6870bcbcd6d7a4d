"""TEST INFRASTRUCTURE -- records the state_dict keys / shapes of every reference processor built with default arguments
(executing the reference through oracle/ref_loader.py) into tests/golden/state_dict_keys.json, so that the CPU test
tests/test_host_logic_cpu.py::test_state_dict_keys_match_the_reference can check, without /root/reference, that upstream
checkpoints load into the drop-in modules with strict=True.  Run:  python -m oracle.make_state_dict_keys"""
import importlib
import inspect
import json
import os
import sys

import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle.ref_loader import load_reference

    load_reference()
    ref = importlib.import_module("grafx.processors")
    out = {}
    for n in sorted(dir(ref)):
        cls = getattr(ref, n)
        if not (inspect.isclass(cls) and issubclass(cls, nn.Module)):
            continue
        try:
            m = cls()
        except Exception as e:  # containers need arguments; FIRFilter cannot be constructed upstream
            out[n] = {"error": type(e).__name__}
            continue
        out[n] = {k: [list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()}
    path = os.path.join(HERE, "..", "tests", "golden", "state_dict_keys.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path, len(out), "classes")


if __name__ == "__main__":
    main()
