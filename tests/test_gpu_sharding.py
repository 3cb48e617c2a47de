"""GPU, world size 2: `sharding.write_sharded` with the real `_native.Codec` on every rank (ranks take GPU
rank % device_count, so this also runs on a one-GPU box).  The two ranks write ONE .cbin/.ch; it must equal, byte for
byte, what a single Writer produces (deterministic, chunk-independent encoder), open in the GPU Reader, and every
chunk must inflate with CPython's zlib — the reference Reader's decoder — to the oracle's transform bytes."""
import hashlib
import json
import os
import socket
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NS, NC, NCHUNK = 3000, 385, 7


def _worker(rank, world, port, tmp):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import mtscomp_b200 as M
    from mtscomp_b200 import _native, sharding
    M.CONFIG_PATH = os.path.join(tmp, '.mtscomp')
    dev = rank % max(_native.load_library().mtsb_device_count(), 1)
    sharding.write_sharded(os.path.join(tmp, 'data.bin'), os.path.join(tmp, 'sharded.cbin'),
                           os.path.join(tmp, 'sharded.ch'), rank, world, sample_rate=float(NS), n_channels=NC,
                           dtype=np.int16, device=dev)
    dist.destroy_process_group()


def test_two_gpu_ranks_write_one_cbin(tmp_path, codec):
    import torch.multiprocessing as mp
    import mtscomp_b200 as M
    from mtscomp_b200 import synth
    from oracle import codec as ora
    M.CONFIG_PATH = tmp_path / '.mtscomp'
    data = np.concatenate([synth.ap_chunk(NS, NC, seed=40 + i) for i in range(NCHUNK)])[:NS * NCHUNK - 1234]
    data.tofile(tmp_path / 'data.bin')
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    # the sequential Writer of this process gives the same file
    M.compress(tmp_path / 'data.bin', tmp_path / 'single.cbin', tmp_path / 'single.ch', sample_rate=float(NS),
               n_channels=NC, dtype=np.int16, quiet=True)
    blob = (tmp_path / 'sharded.cbin').read_bytes()
    assert blob == (tmp_path / 'single.cbin').read_bytes()
    meta, single = json.loads((tmp_path / 'sharded.ch').read_text()), json.loads((tmp_path / 'single.ch').read_text())
    assert meta == single
    assert meta['sha1_compressed'] == hashlib.sha1(blob).hexdigest()
    assert meta['sha1_uncompressed'] == hashlib.sha1(data.tobytes()).hexdigest()
    r = M.decompress(tmp_path / 'sharded.cbin', tmp_path / 'sharded.ch')
    assert np.array_equal(r[:], data)
    r.close()
    b, o = meta['chunk_bounds'], meta['chunk_offsets']
    assert len(b) == NCHUNK + 1
    for i in range(NCHUNK):
        assert zlib.decompress(blob[o[i]:o[i + 1]]) == ora.transform_chunk(data[b[i]:b[i + 1]])
