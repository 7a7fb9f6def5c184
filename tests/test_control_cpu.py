"""The batched ReadUntil loop's host logic on the CPU, with the classifier replaced by canned decisions: CSV rows
(format of riser/control.py:148-153), unblock / stop_receiving calls and their order (control.py:100-106), the
minute tally (control.py:110-117), the stop messages (control.py:118-124)."""
import logging

import numpy as np

from riser_b200 import control as ctl_mod
from riser_b200 import sim
from riser_b200.pipeline import ACCEPT, REJECT, NO_DECISION, SKIPPED, TRY_AGAIN


class _Result:
    pass


class _FakeClassifier:
    """decision of channel c at poll k = PLAN[k][c - 1]"""
    PLAN = [[SKIPPED, SKIPPED, TRY_AGAIN, REJECT], [ACCEPT, SKIPPED, NO_DECISION]]

    def __init__(self, models, processor):
        self.poll = 0
        self.seen = []

    def classify_batch(self, signals, read_ids, cache, threshold, mode):
        codes = self.PLAN[self.poll][:len(signals)]
        self.poll += 1
        self.seen.append(list(read_ids))
        res = _Result()
        res.decisions = np.array(codes, dtype=np.uint8)
        res.sig_len = np.array([len(s) - 1 for s in signals], dtype=np.int32)
        res.p_on = np.array([[0.25 + 0.125 * i, 0.5] for i in range(len(signals))], dtype=np.float32)
        return res


class _Model:
    def __init__(self, target):
        self.target = target


def test_loop_rows_calls_and_messages(tmp_path, monkeypatch, caplog):
    monkeypatch.setattr(ctl_mod, "BatchedClassifier", _FakeClassifier)
    reads = [(f"read-{c}", np.arange(4000, dtype=np.int16)) for c in range(4)]
    client = sim.SimClient(reads, chunk=1000, n_polls=2, first_len=1000)
    log = logging.getLogger("ctl-test")
    control = ctl_mod.SequencerControl(client, [_Model("mRNA"), _Model("globin")], None, log, str(tmp_path / "out"))
    with caplog.at_level(logging.INFO, logger="ctl-test"):
        control.start()
        control.target("deplete", 1, 0.9)
        control.finish()
    lines = (tmp_path / "out.csv").read_text().splitlines()
    assert lines[0] == "batch_start,read_id,channel,sig_length,models,prob_targets,threshold,mode,decision"
    rows = [ln.split(",") for ln in lines[1:]]
    assert [r[1:] for r in rows] == [
        ["read-2", "3", "999", "mRNA;globin", "0.5;0.5", "0.9", "deplete", "try_again"],
        ["read-3", "4", "999", "mRNA;globin", "0.625;0.5", "0.9", "deplete", "reject"],
        # poll 2: channel 4 was finished after poll 1, the batch is channels 1, 2, 3
        ["read-0", "1", "1999", "mRNA;globin", "0.25;0.5", "0.9", "deplete", "accept"],
        ["read-2", "3", "1999", "mRNA;globin", "0.5;0.5", "0.9", "deplete", "no_decision"],
    ]
    assert all(r[0].isdigit() for r in rows)
    assert client.unblocked == [(4, 4)]                              # (channel, read.number): minknow-api <= v5
    assert client.finished == [(4, 4), (1, 1), (3, 3)]               # rejects, then accepts, then undecided
    assert control.classifier.seen == [["read-0", "read-1", "read-2", "read-3"], ["read-0", "read-1", "read-2"]]
    assert list(control.batch_sizes) == [4, 3] and len(control.batch_latencies) == 2
    assert client.messages[0].startswith("The sequencing run is being controlled by RISER")
    assert client.messages[-1] == "RISER has stopped running."
    text = caplog.text
    assert "Live read stream started." in text and "Client has stopped." in text
    assert "Client reset and live read stream ended." in text and "timed out" not in text


def test_minute_tally_and_read_ids_without_number(tmp_path, monkeypatch, caplog):
    monkeypatch.setattr(ctl_mod, "BatchedClassifier", _FakeClassifier)
    clock = {"t": 1000.0}
    monkeypatch.setattr(ctl_mod.time, "monotonic", lambda: clock["t"])
    reads = [(f"r{c}", np.arange(3000, dtype=np.int16)) for c in range(4)]
    client = sim.SimClient(reads, chunk=1000, n_polls=2, first_len=1000, with_number=False)
    orig = client.get_read_batch

    def slow_batch():
        clock["t"] += 61.0                                          # every poll starts a minute after the last
        return orig()
    client.get_read_batch = slow_batch
    log = logging.getLogger("ctl-test2")
    control = ctl_mod.SequencerControl(client, [_Model("mRNA")], None, log, str(tmp_path / "o"))
    with caplog.at_level(logging.INFO, logger="ctl-test2"):
        control.start()
        control.target("enrich", 1, 0.9)
    # poll 1 starts at t = 1000 (not yet past 1060); poll 2 starts at 1061 and reports both polls' counts
    assert caplog.text.count("In the last minute") == 1
    assert "In the last minute 4 signals were assessed, 1 were accepted and 1 were rejected" in caplog.text
    assert client.unblocked == [(4, "r3")] and client.finished == [(4, "r3"), (1, "r0"), (3, "r2")]   # minknow-api >= v6


def test_idle_polls_keep_no_state_and_history_is_bounded(tmp_path, monkeypatch):
    """ReadUntil's get_read_chunks does not block: an idle client is polled tens of thousands of times a second for
    days.  Empty polls must not touch the classifier or grow anything; the latency history is a bounded window; the
    client still sees the reference's two calls with empty lists (control.py:100-106)."""
    class _Idle(sim.SimClient):
        def __init__(self, n_polls):
            super().__init__([], chunk=1000, n_polls=n_polls)
            self.calls = 0

        def get_read_batch(self):
            self.poll += 1
            return []

        def reject_reads(self, reads, unblock_duration):
            assert reads == []
            self.calls += 1

        def finish_processing_reads(self, reads):
            assert reads == []
            self.calls += 1

    class _Never:
        def __init__(self, *a, **k):
            pass

        def classify_batch(self, *a, **k):
            raise AssertionError("an empty poll reached the classifier")

    monkeypatch.setattr(ctl_mod, "BatchedClassifier", _Never)
    client = _Idle(5000)
    control = ctl_mod.SequencerControl(client, [_Model("mRNA")], None, logging.getLogger("idle"), str(tmp_path / "idle"))
    control.start()
    control.target("deplete", 1, 0.9)
    assert client.calls == 2 * 5000
    assert len(control.batch_latencies) == 0 and len(control.batch_sizes) == 0
    assert control.batch_latencies.maxlen == ctl_mod.LATENCY_HISTORY == control.batch_sizes.maxlen
