"""Extract the facts the oracle relies on from the reference's checked-in TensorBoard graphs and write tests/golden/graph_facts.json.

Source: /root/reference/src/~/reacher/data/viz/1/events.out.tfevents.* (12 files written by `tf.summary.FileWriter(...).add_graph`,
lstm_train.py:89-90; the graphs are the two-headed LSTM experiment of backup/student_rollout.py:130-170 at debug size -- 2 unrolled
steps, batch 2, 1 LSTM unit -- next to the full-size frozen teacher `pi/*` and the Adam optimiser).  Run in the build container only
(the reference tree does not exist on the GPU box):

    python tests/golden/make_graph_facts.py

tests/test_graph_facts.py then asserts that oracle/nn_np.py, oracle/lstm_np.py, oracle/lstm2_np.py and AdamTF agree with these facts.
"""
import glob
import json
import os

from tensorboard.backend.event_processing.event_file_loader import EventFileLoader
from tensorboard.compat.proto import graph_pb2
from tensorboard.util import tensor_util

SRC = "/root/reference/src/~/reacher/data/viz/1"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "graph_facts.json")


def load_graph(path):
    for ev in EventFileLoader(path).Load():
        if ev.HasField("graph_def"):
            g = graph_pb2.GraphDef()
            g.ParseFromString(ev.graph_def)
            return g
    raise RuntimeError("no graph_def in " + path)


def facts_of(g):
    N = {n.name: n for n in g.node}

    def const(name):
        v = tensor_util.make_ndarray(N[name].attr["value"].tensor)
        return v.tolist() if v.ndim else v.item()

    def shape(name):
        return [d.size for d in N[name].attr["shape"].shape.dim]

    def ins(name):
        return [i.split(":")[0].lstrip("^") for i in N[name].input]

    def op(name):
        return N[name].op

    # ---- teacher: obfilter -> clip -> fc1 tanh fc2 tanh final ; pdparam = concat([mean, 0 * mean + logstd]) ----------------
    t = {}
    assert op("pi/obfilter/Maximum") == "Maximum" and "pi/obfilter/sub" in ins("pi/obfilter/Maximum")
    assert ins("pi/obfilter/sub") == ["pi/obfilter/ToFloat_1", "pi/obfilter/Square"]          # E[x^2] - mean^2
    assert ins("pi/obfilter/truediv") == ["pi/obfilter/runningsum/read", "pi/obfilter/count/read"]
    assert op("pi/obfilter/Sqrt") == "Sqrt" and ins("pi/obfilter/Sqrt") == ["pi/obfilter/Maximum"]
    t["obfilter_var_floor"] = const("pi/obfilter/Maximum/y")
    assert ins("pi/vf/sub") == ["pi/ob", "pi/obfilter/ToFloat"] and ins("pi/vf/truediv") == ["pi/vf/sub", "pi/obfilter/Sqrt"]
    assert op("pi/vf/clip_by_value/Minimum") == "Minimum" and op("pi/vf/clip_by_value") == "Maximum"
    t["clip"] = [const("pi/vf/clip_by_value/y"), const("pi/vf/clip_by_value/Minimum/y")]
    chain, x = [], "pi/vf/clip_by_value"
    for layer, act in (("fc1", "pi/pol/Tanh"), ("fc2", "pi/pol/Tanh_1"), ("final", None)):
        mm, ba = "pi/pol/%s/MatMul" % layer, "pi/pol/%s/BiasAdd" % layer
        assert ins(mm) == [x, "pi/pol/%s/kernel/read" % layer] and ins(ba) == [mm, "pi/pol/%s/bias/read" % layer]
        assert not N[mm].attr["transpose_a"].b and not N[mm].attr["transpose_b"].b
        chain.append(dict(layer=layer, kernel=shape("pi/pol/%s/kernel" % layer), bias=shape("pi/pol/%s/bias" % layer),
                          activation="tanh" if act else "linear"))
        if act:
            assert op(act) == "Tanh" and ins(act) == [ba]
            x = act
    t["layers"] = chain
    t["logstd_shape"] = shape("pi/pol/logstd")
    assert ins("pi/pol/mul") == ["pi/pol/final/BiasAdd", "pi/pol/mul/y"] and const("pi/pol/mul/y") == 0.0
    assert ins("pi/pol/add") == ["pi/pol/mul", "pi/pol/logstd/read"]
    assert ins("pi/pol/concat")[:2] == ["pi/pol/final/BiasAdd", "pi/pol/add"] and const("pi/pol/concat/axis") == 1
    t["pdparam"] = "concat([mean, 0 * mean + logstd], axis=1)"
    t["obfilter_shapes"] = dict(runningsum=shape("pi/obfilter/runningsum"), runningsumsq=shape("pi/obfilter/runningsumsq"), count=shape("pi/obfilter/count"))
    # ---- Adam --------------------------------------------------------------------------------------------------------------------
    apply_nodes = [n for n in g.node if n.op == "ApplyAdam"]
    a = dict(learning_rate=const("adam/Adam/learning_rate"), beta1=const("adam/Adam/beta1"), beta2=const("adam/Adam/beta2"),
             epsilon=const("adam/Adam/epsilon"), use_nesterov=sorted({bool(n.attr["use_nesterov"].b) for n in apply_nodes}),
             n_apply_nodes=len(apply_nodes))
    # ---- LSTM cell: z = [x, m_prev] W + b ; i, j, f, o = split(z, 4) ; c = sig(f + 1) c_prev + sig(i) tanh(j) ; m = sig(o) tanh(c) -------
    c = "LSTM/unique_lstm_cell/"
    assert ins(c + "MatMul") == [c + "concat", c + "kernel/read"] and ins(c + "BiasAdd") == [c + "MatMul", c + "bias/read"]
    assert ins(c + "split") == [c + "split/split_dim", c + "BiasAdd"] and const(c + "Const") == 4
    assert N[c + "add"].input[0] == c + "split:2" and N[c + "Sigmoid_1"].input[0] == c + "split" and N[c + "Tanh"].input[0] == c + "split:1"
    assert N[c + "Sigmoid_2"].input[0] == c + "split:3"
    assert ins(c + "mul") == [c + "Sigmoid", "LSTM/cm_state/control_dependency"]                       # sig(f + 1) * c_prev
    assert ins(c + "mul_1") == [c + "Sigmoid_1", c + "Tanh"] and ins(c + "add_1") == [c + "mul", c + "mul_1"]
    assert ins(c + "Tanh_1") == [c + "add_1"] and ins(c + "mul_2") == [c + "Sigmoid_2", c + "Tanh_1"]
    # second unrolled step: the SAME kernel, fed with the first step's (c, m): the state is carried through the unroll
    assert ins(c + "MatMul_1") == [c + "concat_1", c + "kernel/read"]
    assert ins(c + "concat_1")[1] == c + "mul_2" and ins(c + "mul_3") == [c + "Sigmoid_3", c + "add_1"]
    kshape = shape(c + "kernel")
    units = kshape[1] // 4
    steps = len([n for n in g.node if n.op == "MatMul" and n.name.startswith(c + "MatMul")])
    heads = {}
    for n in g.node:
        if n.op == "VariableV2" and n.name.startswith("LSTM/") and "Adam" not in n.name and "unique_lstm_cell" not in n.name:
            heads[n.name[len("LSTM/"):]] = shape(n.name)
    l = dict(forget_bias=const(c + "add/y"), gate_order=["i", "j", "f", "o"], kernel_shape=kshape, units=units,
             input_dim=kshape[0] - units, concat_order=["input", "m_prev"], state_carried_through_unroll=True, cell_shared_by_steps=True,
             unrolled_steps=steps, heads_unshared_per_step=True, head_variables=heads)
    # ---- losses -------------------------------------------------------------------------------------------------------------------
    kl_name = [n.name for n in g.node if n.op == "Sum" and "kl_loss" in n.name and "gradients" not in n.name][0]
    k = dict(node=kl_name, reduction_indices=const(kl_name + "/reduction_indices"), reduce="sum")
    return dict(teacher=t, adam=a, lstm=l, kl=k)


def main():
    files = sorted(glob.glob(os.path.join(SRC, "events.out.tfevents.*")))
    assert files, "reference tfevents not found (run this in the build container)"
    per_file = [facts_of(load_graph(f)) for f in files]
    # the 12 graphs are two sizes of the same experiment: the facts the oracle uses must be common to all of them
    common = per_file[-1]
    for pf in per_file:
        assert pf["teacher"] == common["teacher"] and pf["adam"] == common["adam"] and pf["kl"]["reduction_indices"] == common["kl"]["reduction_indices"]
        for key in ("forget_bias", "gate_order", "concat_order", "state_carried_through_unroll", "heads_unshared_per_step"):
            assert pf["lstm"][key] == common["lstm"][key]
    out = dict(source=[os.path.relpath(f, "/root/reference") for f in files], written_by="lstm_train.py:89-90 (tf.summary.FileWriter.add_graph)",
               **common, lstm_variants=sorted({json.dumps(dict(kernel_shape=pf["lstm"]["kernel_shape"], heads=sorted(pf["lstm"]["head_variables"]))) for pf in per_file}))
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
