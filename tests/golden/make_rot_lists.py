"""Rotation index lists of the checked-in emitted ResNets (the static initialiser of
Get_context_params in fhe-cmplr/rtlib/ant/dataset/<model>.onnx.inc) -> tests/emitted/<model>.rots.json.
Run where /root/reference exists; the JSON files travel with the repo."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DS = "/root/reference/fhe-cmplr/rtlib/ant/dataset"
for m in ("resnet20_cifar10_pre", "resnet32_cifar100_pre", "resnet56_cifar10_pre", "resnet110_cifar10_train"):
    src = open(os.path.join(DS, m + ".onnx.inc")).read()
    body = src[src.index("CKKS_PARAMS* Get_context_params()"):]
    body = body[:body.index("return")]
    head = re.search(r"LIB_ANT,\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),", body)
    N, sec, depth, q0, sf, parts, hw, n = (int(x) for x in head.groups())
    rots = [int(x) for x in re.findall(r"-?\d+", body[body.index("{", head.end()):])]
    assert len(rots) == n, (m, len(rots), n)
    json.dump({"model": m, "N": N, "mul_depth": depth, "first_mod_size": q0, "scaling_mod_size": sf,
               "num_q_parts": parts, "hamming_weight": hw, "rot_idxs": rots},
              open(os.path.join(ROOT, "tests", "emitted", m + ".rots.json"), "w"))
    print(m, N, depth, q0, sf, parts, hw, n)
