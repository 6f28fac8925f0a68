"""GPU x2: the CUDA-IPC / NVLink wire.  One process per GPU (the deployment model); the sender's
k_send stores frames straight into the ring in the OTHER GPU's HBM, the receiver's k_recv returns
credit with a 16-byte store back over NVLink, Disconnect writes peer_exit the same way.  Delivered
bytes must equal the sent pattern, the ring must be all-zero once drained, cursors consistent."""
import json
import os
import subprocess
import sys
import tempfile

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(ring_kb, msg, n_msgs, conns):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box (run with gpurun --gpus 2)")
    with tempfile.TemporaryDirectory() as d:
        procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "ipc_wire_worker.py"), role, str(dev), d,
                                   str(ring_kb), str(msg), str(n_msgs), str(conns)],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                 for role, dev in (("server", 1), ("client", 0))]
        outs = [p.communicate(timeout=500)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(outs)
        return [json.load(open(os.path.join(d, r + ".json"))) for r in ("client", "server")]


def test_stream_across_two_gpus():
    cli, srv = _run(4096, 1 << 20, 6, 4)          # 6 x 1 MiB on 4 connections, 4 MiB rings: wraps the ring
    assert srv["ok"] and srv["ring_empty"] and srv["half_closed"]
    assert srv["state"]["remain"] == 0 and srv["state"]["head"] == srv["state"]["moving_head"]


def test_message_larger_than_the_ring_needs_credit_over_nvlink():
    cli, srv = _run(256, 3 << 20, 2, 1)           # 3 MiB messages through a 256 KiB ring
    assert srv["ok"] and srv["ring_empty"] and srv["half_closed"]


def test_dead_peer_process_is_detected():
    """The liveness leg of get_status (pair.cc:358-372): a peer process that dies without Disconnect leaves no
    peer_exit write behind; the survivor's status probe (every 500 ms) turns the pair HALF_CLOSED.  Both processes
    share GPU 0, so this also runs on a one-GPU box."""
    with tempfile.TemporaryDirectory() as d:
        env = dict(os.environ, IPC_WIRE_MODE="death")
        procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "ipc_wire_worker.py"), role, "0", d, "64", "1024", "1", "1"],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
                 for role in ("server", "client")]
        outs = [p.communicate(timeout=200)[0] for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(outs)
        res = json.load(open(os.path.join(d, "client.json")))
        assert res["half_closed"] and res["seconds"] < 10, res
