"""GPU tier, needs >= 2 visible GPUs (skipped otherwise): the row-sharded pipelines over NCCL
must equal the single-index result bit for bit (tools/multi_gpu_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_pipelines_over_nccl_equal_single_index():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29517",
                          os.path.join(ROOT, "tools", "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MISMATCH" not in out.stdout


def test_vectorstore_over_two_real_devices_equals_one_device():
    """``VectorStore({"devices": [0, 1]})``: one process, one host thread, the collection dealt out over two
    GPUs, NVLink peer stores into each other's receive buffers, one merge on the first device."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, ROOT)
    from compressed_rag_suite_b200.rag import Chunk, VectorStore
    rng = np.random.default_rng(51)
    n, dim = 20000, 384
    x = rng.standard_normal((n, dim)).astype(np.float32)
    x[15000] = x[3]
    chunks = [Chunk(f"text {i}", f"chunk_{i}", 0, 6, page_number=i % 13) for i in range(n)]
    devs = list(range(min(torch.cuda.device_count(), 8)))
    one = VectorStore({"collection_name": "one"})
    many = VectorStore({"collection_name": "many", "devices": devs})
    for lo, hi in [(0, 12000), (12000, n)]:
        one.create_index(chunks[lo:hi], x[lo:hi])
        many.create_index(chunks[lo:hi], torch.from_numpy(x[lo:hi]).cuda())       # device tensors, peer-copied to their shard
    for i in (3, 500, 19999):
        for where in (None, {"page_number": {"$lt": 4}}):
            assert many.search(x[i], top_k=10, where=where) == one.search(x[i], top_k=10, where=where)
    assert many.search(x[3], top_k=2)["ids"] == [["chunk_3", "chunk_15000"]]
    assert not any(t for t, _ in many.collection.index.exchange_status())
