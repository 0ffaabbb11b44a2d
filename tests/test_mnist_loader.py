"""MnistDataLoader (include/MnistDataLoader.hpp; reference include/MnistDataLoader.hpp:10-51, src/MnistDataLoader.cpp:9-137)
instantiates and streams IDX files through DataSet with the reference's chunk protocol.  CPU only: host code."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import PKG_NAME, REPO

import importlib

DRIVER = os.path.join(REPO, "tests", "cpp", "mnist_loader_test")


def write_idx(d, images, labels):
    with open(os.path.join(d, "train-images-idx3-ubyte"), "wb") as f:
        f.write(struct.pack(">4I", 0x803, len(images), 28, 28))
        f.write(images.tobytes())
    with open(os.path.join(d, "train-labels-idx1-ubyte"), "wb") as f:
        f.write(struct.pack(">2I", 0x801, len(labels)))
        f.write(labels.tobytes())


@pytest.mark.parametrize("n,max_load", [(25, 10), (25, 0), (30, 30), (7, 100)])
def test_mnist_loader_streams_idx_files(tmp_path, n, max_load):
    importlib.import_module(PKG_NAME).build()  # no-op when everything is up to date
    assert os.path.exists(DRIVER)
    rng = np.random.default_rng(n)
    images = rng.integers(0, 256, (n, 784), dtype=np.uint8)
    labels = rng.integers(0, 10, n, dtype=np.uint8)
    write_idx(str(tmp_path), images, labels)
    out = subprocess.run([DRIVER, str(tmp_path), str(max_load)], capture_output=True, text=True, check=True, timeout=120).stdout.splitlines()
    assert out[0] == "depth 794 name0 0x0 name783 27x27 name784 label:0 name793 label:9 cols 794"
    rows = np.concatenate([images.astype(np.float64), np.eye(10)[labels]], axis=1)
    k = np.arange(1, 795)
    step = max_load if max_load else n
    pos, li = 0, 0
    while True:
        chunk = rows[pos:pos + step]
        wsum = float(sum((((r + 1) * k) % 1000 * chunk[r]).sum() for r in range(len(chunk))))
        # the stream wraps when a load comes back empty (src/MnistDataLoader.cpp:49-53); a short dataset is never ">= 60000"
        pos = pos + len(chunk) if len(chunk) else 0
        at_start = 1 if pos == 0 else 0
        assert out[1 + li] == f"load {li} rows {len(chunk)} atStart {at_start} sum {chunk.sum():.1f} wsum {wsum:.1f} valid 1", out[1 + li]
        li += 1
        if at_start:
            break
    assert out[1 + li] == f"next pass rows {min(step, n)}"
    assert out[2 + li] == f"preview {min(3, n)} weight 1.0 binary 0 continuous 1 spec 794"
