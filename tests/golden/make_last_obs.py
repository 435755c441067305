"""Extract the observation batches the reference's own training runs left behind.

Every saved SB3 policy under /root/reference/train_improved*/models/*.zip embeds `_last_obs`, the
(64, 107) float32 observation batch of the run's last step (training preset, 64 envs, produced by the
unmodified plantos_env.py under SB3 2.7.0 on the authors' machine).  They are the only numeric
artefacts of the hot path the reference ships, so they are committed as a value fixture:
tests/test_oracle_golden.py checks the float tables and the oracle's observations against them.

    python tests/golden/make_last_obs.py        # needs /root/reference; writes ref_last_obs.npz
"""
import base64, glob, json, os, zipfile

import cloudpickle
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = {}
    for path in sorted(glob.glob("/root/reference/train_improved*/models/*.zip")):
        data = json.loads(zipfile.ZipFile(path).read("data"))
        obs = cloudpickle.loads(base64.b64decode(data["_last_obs"][":serialized:"]))
        assert obs.shape == (64, 107) and obs.dtype == np.float32
        key = os.path.relpath(path, "/root/reference").replace("/", "__").replace(".zip", "")
        out[key] = obs
    np.savez_compressed(os.path.join(HERE, "ref_last_obs.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
