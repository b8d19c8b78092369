"""BASELINE config 4 shape on CPU: the OpenAI server on top of a multi-process chain pipeline (torchrun, gloo):
rank 0 = HTTP + first stage, ranks 1.. = stage workers; `/v1/chat/completions` answers == single-process answers."""
import http.client
import json
import os
import signal
import socket
import subprocess
import sys
import time

import pytest
import torch

from helpers import TINY_LLAMA
from mlx_sharding_b200.utils.checkpoint import write_synthetic_checkpoint

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _post(port, body):
    c = http.client.HTTPConnection("127.0.0.1", port, timeout=120)
    c.request("POST", "/v1/chat/completions", json.dumps(body), {"Content-Type": "application/json"})
    r = c.getresponse()
    data = r.read()
    c.close()
    return r.status, json.loads(data)


def _wait_http(port, proc, timeout=150):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if proc.poll() is not None:
            raise RuntimeError("server exited early")
        try:
            c = http.client.HTTPConnection("127.0.0.1", port, timeout=2)
            c.request("GET", "/health")
            if c.getresponse().status == 200:
                return
        except OSError:
            time.sleep(0.5)
    raise TimeoutError("server did not come up")


@pytest.mark.timeout(400)
def test_chat_completions_over_4_stage_chain(tmp_path):
    ckpt = write_synthetic_checkpoint(str(tmp_path / "tiny"), TINY_LLAMA, dtype=torch.float32)
    env = dict(os.environ, PYTHONPATH=REPO, OMP_NUM_THREADS="1")
    body = {"messages": [{"role": "user", "content": "hello pipeline"}], "max_tokens": 8, "temperature": 0}
    answers = []
    for nproc in (1, 4):
        http_port, master_port = _free_port(), _free_port()
        if nproc == 1:
            cmd = [sys.executable, "-m", "shard.openai_api"]
        else:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
                   "127.0.0.1", "--master-port", str(master_port), "-m", "shard.openai_api"]
        cmd += ["--model", ckpt, "--port", str(http_port), "--device", "cpu", "--kv-pages", "64", "--page-size", "16"]
        proc = subprocess.Popen(cmd, cwd=str(tmp_path), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                start_new_session=True)
        try:
            _wait_http(http_port, proc)
            status, j = _post(http_port, body)
            assert status == 200 and j["object"] == "chat.completions"
            # a second, concurrent-capable request keeps working (sequence slots are recycled)
            status2, j2 = _post(http_port, body)
            assert j2["choices"][0]["logprobs"]["tokens"] == j["choices"][0]["logprobs"]["tokens"]
            answers.append(j["choices"][0]["logprobs"]["tokens"])
        finally:
            os.killpg(proc.pid, signal.SIGTERM)  # exact process group we started
            try:
                proc.wait(timeout=20)
            except subprocess.TimeoutExpired:
                os.killpg(proc.pid, signal.SIGKILL)
    assert answers[0] == answers[1] and len(answers[0]) == 8


@pytest.mark.timeout(400)
def test_chat_completions_expert_parallel_group(tmp_path):
    """``mlx-sharding-api --expert-parallel`` under torchrun (CPU, gloo): 2 ranks, experts sharded at load, lockstep group of
    engines; concurrent requests land on different ranks and every answer equals the single-process answer."""
    import concurrent.futures

    from helpers import TINY_DSV2

    ckpt = write_synthetic_checkpoint(str(tmp_path / "dsv2"), TINY_DSV2, dtype=torch.float32, seed=2)
    env = dict(os.environ, PYTHONPATH=REPO, OMP_NUM_THREADS="1")
    bodies = [{"messages": [{"role": "user", "content": f"hello experts {i}"}], "max_tokens": 6 + i, "temperature": 0, "logprobs": 1}
              for i in range(4)]
    answers = []
    for nproc in (1, 2):
        http_port, master_port = _free_port(), _free_port()
        if nproc == 1:
            cmd = [sys.executable, "-m", "shard.openai_api"]
        else:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
                   "127.0.0.1", "--master-port", str(master_port), "-m", "shard.openai_api", "--expert-parallel"]
        cmd += ["--model", ckpt, "--port", str(http_port), "--device", "cpu", "--kv-pages", "64", "--page-size", "16"]
        proc = subprocess.Popen(cmd, cwd=str(tmp_path), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                start_new_session=True)
        try:
            _wait_http(http_port, proc)
            with concurrent.futures.ThreadPoolExecutor(4) as ex:
                res = list(ex.map(lambda b: _post(http_port, b), bodies))
            assert all(st == 200 for st, _ in res)
            answers.append([j["choices"][0]["logprobs"]["tokens"] for _, j in res])
            if nproc == 2:
                c = http.client.HTTPConnection("127.0.0.1", http_port, timeout=10)
                c.request("GET", "/metrics")
                text = c.getresponse().read().decode()
                assert "lockstep_assigned_rank0 2" in text and "lockstep_assigned_rank1 2" in text, text
        finally:
            os.killpg(proc.pid, signal.SIGTERM)
            try:
                proc.wait(timeout=20)
            except subprocess.TimeoutExpired:
                os.killpg(proc.pid, signal.SIGKILL)
    assert answers[0] == answers[1] and [len(a) for a in answers[0]] == [6, 7, 8, 9]


@pytest.mark.timeout(300)
def test_manual_chain_launch_without_torchrun(tmp_path):
    """The native chain can also be assembled by hand, reference-style: ``mlx-sharding-server --rank 1 --world-size 2`` for the
    second stage and ``mlx-sharding-api`` with RANK / WORLD_SIZE in its environment as the front end (default KV pools match)."""
    ckpt = write_synthetic_checkpoint(str(tmp_path / "tiny"), TINY_LLAMA, dtype=torch.float32)
    env = dict(os.environ, PYTHONPATH=REPO, OMP_NUM_THREADS="1")
    http_port, master_port = _free_port(), _free_port()
    body = {"messages": [{"role": "user", "content": "manual chain"}], "max_tokens": 6, "temperature": 0, "logprobs": 1}
    solo = subprocess.Popen([sys.executable, "-m", "shard.openai_api", "--model", ckpt, "--port", str(http_port), "--device", "cpu"],
                            cwd=str(tmp_path), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, start_new_session=True)
    try:
        _wait_http(http_port, solo)
        _, ref = _post(http_port, body)
    finally:
        os.killpg(solo.pid, signal.SIGTERM)
        solo.wait(timeout=20)
    http_port = _free_port()
    worker = subprocess.Popen([sys.executable, "-m", "shard.main", "--model", ckpt, "--device", "cpu", "--rank", "1", "--world-size", "2",
                               "--master-port", str(master_port)], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, start_new_session=True)
    front = subprocess.Popen([sys.executable, "-m", "shard.openai_api", "--model", ckpt, "--port", str(http_port), "--device", "cpu"],
                             cwd=str(tmp_path), env=dict(env, RANK="0", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(master_port)),
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, start_new_session=True)
    try:
        _wait_http(http_port, front)
        status, got = _post(http_port, body)
        assert status == 200
        assert got["choices"][0]["logprobs"]["tokens"] == ref["choices"][0]["logprobs"]["tokens"]
    finally:
        for p in (front, worker):
            os.killpg(p.pid, signal.SIGTERM)
        for p in (front, worker):
            try:
                p.wait(timeout=20)
            except subprocess.TimeoutExpired:
                os.killpg(p.pid, signal.SIGKILL)
