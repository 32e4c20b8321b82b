"""Bit-exact control-plane parity with the reference scheduler.

tests/golden/control_plane.json was produced by running the UNMODIFIED
reference (/root/reference/tfmesos) under stubs (tests/golden/make_golden.py).
Here the same scripts (tests/golden/scenarios.py) drive tfmesos_b200 and the
recordings must be identical, after dropping the one documented extra
(CUDA_VISIBLE_DEVICES in the TaskInfo environment).
"""
import json
import os
import subprocess
import sys
import tempfile

import pytest

import tfmesos_b200
from tfmesos_b200 import scheduler as sched_mod
from tfmesos_b200 import utils
from tfmesos_b200.utils import AttrDict
from tests.golden import scenarios

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "control_plane.json")))
EXTRA_ENV = {"CUDA_VISIBLE_DEVICES"}


def make_scheduler(jobs, **kw):
    return sched_mod.TFMesosScheduler([sched_mod.Job(**j) for j in jobs],
                                      master="stub", quiet=True, **kw)


def canon(x):
    return json.loads(json.dumps(x, sort_keys=True))


def strip_extra_env(placement):
    for rounds in placement.values():
        for rnd in rounds:
            for ev in rnd["events"]:
                if ev[0] == "launch":
                    for ti in ev[2]:
                        ti["env"] = [kv for kv in ti["env"] if kv[0] not in EXTRA_ENV]
    return placement


def test_placement_matches_reference():
    got = scenarios.run_placement(
        make_scheduler, lambda spec: scenarios.build_offer(AttrDict, spec))
    got = strip_extra_env(canon(got))
    assert sorted(got) == sorted(GOLDEN["placement"])
    for name in sorted(got):
        assert got[name] == GOLDEN["placement"][name], name


def test_gpu_slice_is_exported_as_cuda_visible_devices():
    got = scenarios.run_placement(
        make_scheduler, lambda spec: scenarios.build_offer(AttrDict, spec))
    launch = got["ps2_w4_one_offer_8gpu"][0]["events"][0][2]
    cvd = {ti["name"]: dict(ti["env"]).get("CUDA_VISIBLE_DEVICES") for ti in launch}
    assert cvd == {"/job:ps/task:0": "GPU-0", "/job:ps/task:1": "GPU-1",
                   "/job:worker/task:0": "GPU-2", "/job:worker/task:1": "GPU-3",
                   "/job:worker/task:2": "GPU-4", "/job:worker/task:3": "GPU-5"}
    launch = got["tfrun_s1_w2_gw1"][0]["events"][0][2]
    assert "CUDA_VISIBLE_DEVICES" not in dict(launch[0]["env"])     # -Gs 0: not pinned


def test_cluster_def_config_and_targets_match_reference():
    got = canon(scenarios.run_cluster_def(make_scheduler, D=AttrDict))
    assert got == GOLDEN["cluster_def"]


def test_status_state_machine_matches_reference():
    got = canon(scenarios.run_status(make_scheduler, lambda **kw: AttrDict(**kw)))
    assert got == GOLDEN["status"]


def test_job_normalisation_matches_reference():
    got = canon(scenarios.run_job_normalisation(
        cluster=tfmesos_b200.cluster, Job=sched_mod.Job, sched_module=sched_mod,
        holder_module=tfmesos_b200))
    assert got == GOLDEN["job_normalisation"]


def test_wire_format_matches_reference():
    assert canon(scenarios.run_wire(utils)) == GOLDEN["wire"]


def test_recv_reassembles_frames_larger_than_one_segment():
    import socket
    import threading
    a, b = socket.socketpair()
    big = {"cluster_def": {"worker": ["h%d:%d" % (i, i) for i in range(200000)]}}
    t = threading.Thread(target=utils.send, args=(a, big))
    t.start()
    assert utils.recv(b) == big
    t.join()


def test_replica_mode_end_to_end_matches_reference():
    """cluster() with the real LocalSchedulerDriver: children are spawned through
    tfmesos_b200.server, expand {ps_hosts}... and see the TFMESOS_* env exactly
    as the reference's children did."""
    tmp = tempfile.mkdtemp(prefix="tfmesos_b200_")
    probe = os.path.join(HERE, "golden", "probe_child.py")
    cmd = ("%s %s %s --ps_hosts {ps_hosts} --worker_hosts {worker_hosts} "
           "--job_name {job_name} --worker_index {task_index}"
           % (sys.executable, probe, tmp))
    jobs = [dict(name="ps", num=1, cmd=cmd), dict(name="worker", num=2, cmd=cmd)]
    import time
    hooks = os.path.join(tmp, "hooks.txt")
    extra = {"initializer": "echo init >> %s" % hooks, "finalizer": "echo fin >> %s" % hooks}
    with tfmesos_b200.cluster(jobs, quiet=True, extra_config=extra) as c:
        names = {t.mesos_task_id: (t.job_name, t.task_index, t.addr)
                 for t in c.tasks.values()}
        targets = dict(c.targets)
        assert c.started
        deadline = time.time() + 60
        while not c.finished():
            assert time.time() < deadline
            time.sleep(0.05)
        time.sleep(0.3)
        job_finished = dict(c.job_finished)
    addr_mask = {a: "<%s:%s>" % (j, i) for (j, i, a) in names.values()}
    gold = GOLDEN["replica_mode"]
    assert {k: scenarios.mask_addr(v, addr_mask) for k, v in targets.items()} == gold["targets"]
    assert job_finished["worker"] == gold["job_finished"]["worker"]
    lines = open(hooks).read().split()
    # extra_config initializer / finalizer run once per task (server.py:68-70,106-109)
    assert {"init": lines.count("init"), "fin": lines.count("fin")} == gold["extra_config_hooks"]
    children = {}
    for fn in sorted(os.listdir(tmp)):
        if not fn.endswith(".json"):
            continue
        rec = json.load(open(os.path.join(tmp, fn)))
        rec["argv"] = [scenarios.mask_addr(a, addr_mask) for a in rec["argv"]]
        rec["env"] = {k: scenarios.mask_addr(v, addr_mask) for k, v in rec["env"].items()}
        rec["cwd"] = "<cwd>" if rec["cwd"] == os.getcwd() else rec["cwd"]
        children[fn[:-5]] = rec
    assert children == gold["children"]


def test_prestart_crash_is_retried_then_fatal():
    """Failure policy with real processes: a task that dies before the
    rendez-vous is relaunched, the third failure raises (scheduler.py:404-434)."""
    with pytest.raises(RuntimeError):
        orig = sched_mod.Task.to_task_info

        def broken(self, *a, **k):
            ti = orig(self, *a, **k)
            ti.command.value = "exit 3"
            return ti

        sched_mod.Task.to_task_info = broken
        try:
            with tfmesos_b200.cluster([dict(name="worker", num=1)], quiet=True):
                pass
        finally:
            sched_mod.Task.to_task_info = orig


def test_tfrun_cli_runs_and_expands_arguments():
    tmp = tempfile.mkdtemp(prefix="tfrun_b200_")
    probe = os.path.join(HERE, "golden", "probe_child.py")
    root = os.path.dirname(HERE)
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run(
        [sys.executable, os.path.join(root, "script", "tfrun"), "-w", "2", "-s", "1",
         "--worker-logs", "*", "--", sys.executable, probe, tmp,
         "--job_name", "{job_name}", "--worker_index", "{task_index}",
         "--ps_hosts", "{ps_hosts}", "--worker_hosts", "{worker_hosts}"],
        env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert out.returncode == 0, out.stderr.decode()
    files = sorted(os.listdir(tmp))
    assert files == ["ps_0.json", "worker_0.json", "worker_1.json"]
    text = out.stdout.decode()
    assert "[worker:0] probe child worker_0.json" in text
    assert "[worker:1] probe child worker_1.json" in text


def test_unplaceable_job_fails_loudly_instead_of_waiting_for_more_offers():
    """One box = one offer: asking for more GPUs than it has must raise, not hang
    (on Mesos the reference would keep waiting for further offers)."""
    with pytest.raises(RuntimeError, match="cannot place"):
        with tfmesos_b200.cluster([dict(name="worker", num=2, gpus=4096)], quiet=True):
            pass


def test_local_spawner_pins_each_gpu_task_to_its_slice(monkeypatch, tmp_path):
    """The local driver offers the GPUs named by CUDA_VISIBLE_DEVICES as a SET;
    first-fit hands ps tasks (gpus=0) nothing and each worker one ordinal, which
    the child sees as ITS CUDA_VISIBLE_DEVICES (no GPU needed to check this)."""
    import time
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "4,5,6")
    child = tmp_path / "child.py"
    child.write_text(
        "import json, os, sys\n"
        "open(os.path.join(sys.argv[1], os.environ['TFMESOS_JOB_NAME'] + '_' +\n"
        "     os.environ['TFMESOS_TASK_INDEX'] + '.json'), 'w').write(\n"
        "     json.dumps(os.environ.get('CUDA_VISIBLE_DEVICES')))\n")
    cmd = "%s %s %s" % (sys.executable, child, tmp_path)
    jobs = [dict(name="ps", num=1, cmd=cmd, gpus=0), dict(name="worker", num=3, cmd=cmd, gpus=1)]
    with tfmesos_b200.cluster(jobs, quiet=True) as c:
        slices = {(t.job_name, t.task_index): list(t.gpu_slice) for t in c.tasks.values()}
        deadline = time.time() + 60
        while not c.finished():
            assert time.time() < deadline
            time.sleep(0.05)
        time.sleep(0.3)
    assert slices == {("ps", 0): [], ("worker", 0): ["4"], ("worker", 1): ["5"],
                      ("worker", 2): ["6"]}
    seen = {fn[:-5]: json.load(open(tmp_path / fn)) for fn in os.listdir(tmp_path)
            if fn.endswith(".json")}
    assert seen == {"ps_0": "4,5,6", "worker_0": "4", "worker_1": "5", "worker_2": "6"}


def test_local_spawner_refuses_more_gpu_tasks_than_gpus(monkeypatch):
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "0")
    with pytest.raises(RuntimeError, match="cannot place worker:1"):
        with tfmesos_b200.cluster([dict(name="worker", num=2, gpus=1, cmd="true")], quiet=True):
            pass
