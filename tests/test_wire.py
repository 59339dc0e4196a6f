"""Wire-compatible `Slave` gRPC service (proto.proto:37-70): message encoding checked against hand-computed protobuf
bytes, and the service round trip over a real localhost channel with a stand-in device context (CPU)."""
import numpy as np
import pytest

from distributed_sgd_b200.core import wire


class FakeCtx:
    dim = 6

    def __init__(self):
        self.calls = []

    def forward(self, idx, w):
        self.calls.append(("forward", idx.tolist(), w.tolist()))
        return -np.sign(w[idx % self.dim])

    def gradient(self, idx, w):
        self.calls.append(("gradient", idx.tolist(), w.tolist()))
        if len(idx) == 0:
            raise ValueError("Cannot sum an empty list of vectors")
        g = np.zeros(self.dim)
        g[[1, 4]] = [0.5, -2.0]
        return g

    def start_async(self, w, idx, batch, lr, **kw):
        self.calls.append(("start_async", idx.tolist(), batch, lr))

    def stop_async(self):
        self.calls.append(("stop_async",))

    def update_grad(self, idx, val):
        self.calls.append(("update_grad", idx.tolist(), val.tolist()))


def test_sparse_message_bytes_match_the_protobuf_wire_format():
    M = wire.Messages()
    sp = M.Sparse(size=5)
    sp.map[3] = 1.5
    # field 1 (map entry, LEN): 0a 0b | key: 08 03 | value: 11 <1.5 as little-endian double> ; field 2 (size): 10 05
    assert sp.SerializeToString() == bytes.fromhex("0a0b080311000000000000f83f1005")
    req = M.GradientRequest(samples=[1, 2, 300])
    assert req.SerializeToString() == bytes.fromhex("1204" + "0102ac02")          # packed int32 (proto.proto:62)
    back = M.Sparse.FromString(bytes.fromhex("0a0b080311000000000000f83f1005"))
    assert dict(back.map) == {3: 1.5} and back.size == 5


def test_vec_mapping_uses_the_references_one_based_keys():
    M = wire.Messages()
    sp = wire.dense_to_sparse(M, np.array([0.0, 2.0, 0.0, -1.0]), 4)
    assert dict(sp.map) == {2: 2.0, 4: -1.0} and sp.size == 4                      # key == size is legal (quirk Q11)
    np.testing.assert_array_equal(wire.sparse_to_dense(sp, 4), [0.0, 2.0, 0.0, -1.0])
    bad = M.Sparse(size=4)
    bad.map[5] = 1.0
    with pytest.raises(IndexError):
        wire.sparse_to_dense(bad, 4)
    bad0 = M.Sparse(size=4)
    bad0.map[0] = 1.0
    with pytest.raises(IndexError):
        wire.sparse_to_dense(bad0, 4)


def test_service_round_trip_over_localhost():
    import grpc
    ctx = FakeCtx()
    srv = wire.SlaveServicer(ctx, n_train=10, is_async=False)
    server, port = wire.serve_slave(srv, 0)
    try:
        stub = wire.SlaveStub(f"127.0.0.1:{port}")
        M = stub.M
        assert stub.RegisterSlave(M.Node(host="10.0.0.2", port=4001)) == M.Ack()
        assert ("10.0.0.2", 4001) in srv.colleagues
        w = M.Sparse(size=6)
        w.map[2] = 0.5
        w.map[6] = -1.0
        rep = stub.Gradient(M.GradientRequest(weights=w, samples=[0, 3, 9]))
        assert dict(rep.gradUpdate.map) == {2: 0.5, 5: -2.0} and rep.gradUpdate.size == 6
        assert ctx.calls[-1] == ("gradient", [0, 3, 9], [0.0, 0.5, 0.0, 0.0, 0.0, -1.0])
        rep = stub.Forward(M.ForwardRequest(samples=[1, 5], weights=w))
        assert list(rep.predictions) == [-1.0, 1.0]
        with pytest.raises(grpc.RpcError) as e:                                      # data(idx) out of range
            stub.Gradient(M.GradientRequest(weights=w, samples=[10]))
        assert e.value.code() == grpc.StatusCode.UNKNOWN and "IndexError" in e.value.details()
        with pytest.raises(grpc.RpcError) as e:                                      # Vec.sum(empty) throws (Q7)
            stub.Gradient(M.GradientRequest(weights=w, samples=[]))
        assert "empty list" in e.value.details()
        with pytest.raises(grpc.RpcError) as e:                                      # require(async, ...)
            stub.UpdateGrad(M.GradUpdate(gradUpdate=w))
        assert "synchronous mode" in e.value.details()
        stub.UnregisterSlave(M.Node(host="10.0.0.2", port=4001))
        assert not srv.colleagues
        stub.close()
    finally:
        server.stop(0)
    # async flavour
    actx = FakeCtx()
    asrv = wire.SlaveServicer(actx, n_train=10, is_async=True)
    server, port = wire.serve_slave(asrv, 0)
    try:
        stub = wire.SlaveStub(f"127.0.0.1:{port}")
        M = stub.M
        w = M.Sparse(size=6)
        w.map[1] = 0.25
        stub.StartAsync(M.StartAsyncRequest(weights=w, samples=[0, 1, 2], batchSize=1, learningRate=0.5))
        stub.UpdateGrad(M.GradUpdate(gradUpdate=w))
        stub.StopAsync(M.Empty())
        assert [c[0] for c in actx.calls] == ["start_async", "update_grad", "stop_async"]
        assert actx.calls[1] == ("update_grad", [0], [0.25])
        stub.close()
    finally:
        server.stop(0)


class FakeAsyncCtx(FakeCtx):
    """A worker whose outbox (sum of -delta since it was enabled) follows a script, one state per read."""

    def __init__(self, states):
        super().__init__()
        self.states, self.reads, self.enabled = [np.asarray(s, dtype=float) for s in states], 0, False

    def async_outbox_enable(self):
        self.enabled = True
        self.calls.append(("outbox_enable",))

    def async_outbox_read(self):
        s = self.states[min(self.reads, len(self.states) - 1)]
        self.reads += 1
        return s.copy()


def test_async_deltas_reach_colleagues_that_are_not_gpu_peers():
    """core/Slave.scala:104-105: every delta goes to the colleague slaves and to the master.  For colleagues reached over
    gRPC the GPU worker's outbox is forwarded in periods: the receivers must end up with w -= (sum of all deltas)."""
    # outbox states = -(sum of deltas so far): three periods, the last read repeats (nothing new)
    states = [[0, -0.5, 0, 0, 0.25, 0], [0, -0.5, 0, 0, 0.25, 0], [1.0, -0.75, 0, 0, 0.25, 0], [1.0, -0.75, 0, 0, 0.25, -2.0]]
    worker = FakeAsyncCtx(states)
    srv = wire.SlaveServicer(worker, n_train=10, is_async=True, relay_period=3600.0)     # flushed by hand below
    colleague_ctx = FakeCtx()
    colleague = wire.SlaveServicer(colleague_ctx, n_train=10, is_async=True)
    cserver, cport = wire.serve_slave(colleague, 0)
    got_master = []
    mserver, mport = wire.serve_master({"UpdateGrad": lambda req: got_master.append(wire.sparse_to_dense(req.gradUpdate, 6))}, 0)
    srv.master_target = f"127.0.0.1:{mport}"
    M = srv.M
    try:
        srv.RegisterSlave(M.Node(host="127.0.0.1", port=cport))
        w = M.Sparse(size=6)
        srv.StartAsync(M.StartAsyncRequest(weights=w, samples=[0, 1], batchSize=1, learningRate=0.5))
        assert worker.enabled and [c[0] for c in worker.calls] == ["outbox_enable", "start_async"]
        relay = srv.relay
        assert relay.flush() == 2          # {1: 0.5, 4: -0.25} as deltas (w -= delta)
        assert relay.flush() == 0          # nothing new: no message
        assert relay.flush() == 2
        srv.StopAsync(M.Empty())           # final flush forwards the rest
        assert srv.relay is None and relay.sent == 3 and not relay.errors
        sent = [np.zeros(6) for _ in range(3)]
        ups = [c for c in colleague_ctx.calls if c[0] == "update_grad"]
        assert len(ups) == 3 and len(got_master) == 3
        for k, (_, idx, val) in enumerate(ups):
            sent[k][idx] = val
            assert np.array_equal(sent[k], got_master[k])
        assert np.array_equal(sent[0], [0, 0.5, 0, 0, -0.25, 0])
        assert np.allclose(sum(sent), -np.asarray(states[-1]), rtol=0, atol=1e-15)      # telescopes to the whole outbox
    finally:
        cserver.stop(0)
        mserver.stop(0)


def test_async_without_colleagues_needs_no_outbox():
    worker = FakeAsyncCtx([[0] * 6])
    srv = wire.SlaveServicer(worker, n_train=10, is_async=True)
    M = srv.M
    srv.StartAsync(M.StartAsyncRequest(weights=M.Sparse(size=6), samples=[0], batchSize=1, learningRate=0.5))
    assert not worker.enabled and srv.relay is None
    srv.StopAsync(M.Empty())
