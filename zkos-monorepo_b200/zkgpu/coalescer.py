"""Request coalescer: turns concurrent single-proof requests into `zkgpu_prove_batch` calls.

The reference's prover hosts serve one proof per request, each request on its own task/thread
(/root/reference/tee/crates/shielder-prover-tee/src/server.rs:128-195 — one tokio task per vsock client, up to 100
concurrent: tee/crates/shielder-prover-server/src/command_line_args.rs:26).  On a GPU the throughput comes from
batching, so a drop-in host keeps the per-request call (`prove(advice, instance, seed) -> bytes`, blocking, thread-safe)
and batches behind it: requests of the same circuit that arrive within `max_wait_ms` of each other, up to
`max_batch`, go to the device in one call.  SURVEY.md section 8f-2.
"""
import threading
import time

import numpy as np


class ProofCoalescer:
    def __init__(self, prove_batch, max_batch=128, max_wait_ms=2.0):
        """prove_batch(advice (m, A, n, 4), instance (m, p, 4), seeds (m,)) -> list of m proofs"""
        self._prove_batch, self.max_batch, self.max_wait = prove_batch, int(max_batch), max_wait_ms / 1e3
        self._lock = threading.Condition()
        self._queue = []            # (advice, instance, seed, slot)
        self._closed = False
        self.batches = []           # sizes of the batches issued (introspection / tests)
        self._worker = threading.Thread(target=self._run, daemon=True)
        self._worker.start()

    def prove(self, advice, instance, seed):
        """blocking single-proof call, safe from any number of threads"""
        slot = {"event": threading.Event(), "proof": None, "error": None}
        with self._lock:
            if self._closed:
                raise RuntimeError("coalescer is closed")
            self._queue.append((advice, instance, int(seed), slot))
            self._lock.notify_all()
        slot["event"].wait()
        if slot["error"] is not None:
            raise slot["error"]
        return slot["proof"]

    def close(self):
        with self._lock:
            self._closed = True
            self._lock.notify_all()
        self._worker.join()

    def _run(self):
        while True:
            with self._lock:
                while not self._queue and not self._closed:
                    self._lock.wait()
                if not self._queue and self._closed:
                    return
                deadline = time.monotonic() + self.max_wait
                while len(self._queue) < self.max_batch and not self._closed:
                    left = deadline - time.monotonic()
                    if left <= 0:
                        break
                    self._lock.wait(left)
                batch, self._queue = self._queue[: self.max_batch], self._queue[self.max_batch:]
            try:
                proofs = self._prove_batch(np.stack([b[0] for b in batch]), np.stack([b[1] for b in batch]),
                                           np.array([b[2] for b in batch], dtype=np.uint64))
                for b, pr in zip(batch, proofs):
                    b[3]["proof"] = pr
            except Exception as e:   # a failed batch fails its requests, not the service (server.rs:189-190)
                for b in batch:
                    b[3]["error"] = e
            self.batches.append(len(batch))
            for b in batch:
                b[3]["event"].set()
