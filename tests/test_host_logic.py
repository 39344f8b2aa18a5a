"""CPU tests of the host side: the C-ABI library loads and exports every symbol of include/m1b200.h,
the M1 constructor API / config capture, and the traced parameter inventory against the oracle's."""
import ctypes
import os

import pytest
import torch

import m1b200
from m1b200 import _lib
from m1b200.model import losses, optimizers, unets
from oracle import m1_oracle as O

README = dict(filters=(32, 64, 128, 256, 512),
              strides=((1, 1, 1), (1, 2, 2), (1, 2, 2), (2, 2, 2), (2, 2, 2)),
              kernel_sizes=((1, 3, 3), (1, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3)),
              se_reduction=(8, 8, 8, 8, 8), att_sub_samp=((1, 1, 1),) * 4)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libm1b200.so not built (run __graft_entry__.build())"
    h = ctypes.CDLL(_lib.LIB_PATH)
    sigs = _lib.parse_header()
    assert len(sigs) >= 30
    for name in sigs:
        assert hasattr(h, name), name
    h.m1_version.restype = ctypes.c_int
    assert h.m1_version() >= 100
    # no compute without a GPU: creating a context must fail loudly, not fall back
    if not torch.cuda.is_available():
        lib = _lib.lib()
        hnd = ctypes.c_void_p()
        assert lib.m1_ctx_create(0, ctypes.byref(hnd)) != 0
        assert b"no CPU fallback" in lib.m1_last_error()


def test_no_cpu_fallback_in_model():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = unets.networks.M1((4, 16, 16), 3, 2, summary=False, **README)
    with pytest.raises(RuntimeError):
        m.train_step(torch.zeros(1, 4, 16, 16, 3), torch.zeros(1, 4, 16, 16, 2))
    with pytest.raises(RuntimeError):
        m.get_detect_model()(torch.zeros(1, 4, 16, 16, 3))


def test_constructor_defaults_and_assertions():
    # Q7: the class default att_sub_samp has 3 entries, M1Core asserts 4 -> AssertionError
    with pytest.raises(AssertionError):
        unets.networks.M1((4, 16, 16), 3, 2, summary=False)
    with pytest.raises(AssertionError):
        unets.networks.M1((4, 16, 16), 3, 2, summary=False, filters=(8, 16, 32, 64), att_sub_samp=((1, 1, 1),) * 4)
    m = unets.networks.M1((4, 16, 16), 3, 2, summary=False, att_sub_samp=((1, 1, 1),) * 4)
    cfg = m.get_config()
    assert cfg['dropout_rate'] == 0.5 and cfg['dropout_mode'] == 'standard'
    assert cfg['strides'][-1] == (1, 2, 2)                       # Q9: class default last stride
    assert cfg['prob_latent_dims'] == (3, 2, 1) and cfg['name'] == 'UNET-TYPE-M1'
    m2 = unets.networks.M1.from_config(cfg)
    assert m2.params.num_params() == m.params.num_params()


def test_param_inventory_matches_oracle_deterministic():
    m = unets.networks.M1((4, 32, 32), 3, 2, summary=False, **README)
    cfg = O.default_config(**{k: README[k] for k in ('filters', 'strides', 'kernel_sizes')})
    ps = O.ParamStore(dtype=torch.float32)
    O.m1_deterministic(ps, cfg, torch.zeros((1, 4, 32, 32, 3)), O.Noise(0, torch.float32), training=False)
    ours = {n: sp.shape for n, sp in m.params.specs.items()}
    theirs = {n: tuple(t.shape) for n, t in ps.p.items()}
    assert ours == theirs
    assert m.params.num_params(('kernel', 'bias', 'se_kernel', 'se_bias')) == 17517642
    kinds = {n: sp.kind for n, sp in m.params.specs.items()}
    assert kinds == dict(ps.kind)


@pytest.mark.parametrize("ds_mode", ['reference', 'intended'])
def test_param_inventory_matches_oracle_probabilistic(ds_mode):
    m = unets.networks.M1((4, 32, 32), 4, 2, summary=False, dense_skip=True, deep_supervision=True,
                          probabilistic=True, prob_latent_dims=(3, 2, 1, 0), dropout_mode='monte-carlo',
                          ds_in_prob=ds_mode, **README)
    cfg = O.default_config(dense_skip=True, deep_supervision=True, probabilistic=True,
                           prob_latent_dims=(3, 2, 1, 0), dropout_mode='monte-carlo',
                           **{k: README[k] for k in ('filters', 'strides', 'kernel_sizes')})
    ps = O.ParamStore(dtype=torch.float32)
    x, _ = O.synthetic_batch(1, (4, 32, 32))
    O.m1_probabilistic(ps, cfg, x, O.Noise(0, torch.float32), ds_in_prob=ds_mode, with_infer=True)
    ours = {n: sp.shape for n, sp in m.params.specs.items()}
    theirs = {n: tuple(t.shape) for n, t in ps.p.items()}
    assert set(ours) == set(theirs), (sorted(set(ours) ^ set(theirs))[:10])
    assert ours == theirs
    # ~64.0 M trainable parameters in the training graph
    total = m.params.num_params()
    assert 63_900_000 < total < 64_100_000, total
    # flat layout: groups are contiguous, every parameter 64-float aligned
    P = m.params
    for sp in P.specs.values():
        assert sp.offset % 64 == 0
    (k0, k1), (b0, b1), (p0, p1), (f0, f1) = (P.group_range[g] for g in ('kernel', 'bias', 'plain', 'frozen'))
    assert k0 == 0 and k1 == b0 and b1 == p0 and p1 == f0 and f1 == P.total
    # variables the reference creates but no model output depends on (prior sersd0 + logits, Q3): Keras neither
    # trains nor regularises them - they sit in the 'frozen' group, outside every Adam / L2 launch
    frozen = sorted(n for n, sp in P.specs.items() if sp.frozen)
    if ds_mode == 'reference':
        assert frozen and all(n.startswith(('prior/sersd0/', 'prior/logits/')) for n in frozen), frozen[:5]
        assert {n.split('/')[1] for n in frozen} == {'sersd0', 'logits'}
    else:
        assert not frozen


def test_compile_reads_reference_loss_objects():
    m = unets.networks.M1((4, 16, 16), 4, 2, summary=False, probabilistic=True, prob_latent_dims=(3, 2, 1, 0),
                          **README)
    sched = optimizers.CosineDecayRestarts(1e-3, 100, t_mul=2.0, m_mul=1.0, alpha=1e-3)
    m.compile(optimizer=optimizers.Adam(learning_rate=sched, amsgrad=True),
              loss=[losses.Focal(alpha=[0.75, 0.25], gamma=2.0).loss, losses.EvidenceLowerBound().loss],
              loss_weights=[1.0, 10.0])
    assert m.focal.alpha == [0.75, 0.25] and m.loss_weights == [1.0, 10.0] and m.elbo.beta == 1.0
    for step in (0, 1, 50, 99, 100, 150, 299, 300, 1234):
        assert abs(sched(step) - O.cosine_decay_restarts(step, 1e-3, 100, 2.0, 1.0, 1e-3)) < 1e-15
    with pytest.raises(Exception):
        m.compile(loss=losses.Focal(alpha=[1.0, 1.0, 1.0]).loss)


def test_store_config_args_semantics():
    class Toy(unets.modelio.LoadableModel):
        @unets.modelio.store_config_args
        def __init__(self, a, b=2, *, c=3):
            pass

    assert Toy(1).get_config() == {'a': 1, 'b': 2, 'c': 3}
    assert Toy(1, 5, c=7).get_config() == {'a': 1, 'b': 5, 'c': 7}

    class Bare(unets.modelio.LoadableModel):
        pass

    with pytest.raises(RuntimeError):
        Bare().get_config()


def test_dlpack_pointer_export_rejects_cpu_memory():
    t = torch.zeros(4)
    with pytest.raises(_lib.M1Error):
        _lib.ptr(t)
    with pytest.raises(_lib.M1Error):
        _lib.dlpack_device_ptr(t)
    # the capsule's data pointer (+ byte offset) is the tensor's data pointer, also for offset views
    assert _lib.dlpack_device_ptr(t, expect_cuda=False) == t.data_ptr()
    v = torch.arange(64, dtype=torch.float32)[16:32]
    assert _lib.dlpack_device_ptr(v, expect_cuda=False) == v.data_ptr()


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of m1_conv_desc / m1_dropout have the size and field offsets gcc gives the C structs
    of include/m1b200.h (a drifted binding would silently corrupt every launch)."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "m1b200.h"\n'
        'int main(void) {\n'
        '  printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(m1_conv_desc), offsetof(m1_conv_desc, src_c),\n'
        '         offsetof(m1_conv_desc, w_stride_tap), offsetof(m1_conv_desc, w_by_src),\n'
        '         offsetof(m1_conv_desc, accumulate), offsetof(m1_conv_desc, engine), offsetof(m1_conv_desc, tune));\n'
        '  printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(m1_dropout), offsetof(m1_dropout, seed),\n'
        '         offsetof(m1_dropout, stream_id), offsetof(m1_dropout, rate), offsetof(m1_dropout, step),\n'
        '         offsetof(m1_dropout, mask));\n'
        '  printf("%llu\\n", (unsigned long long)M1_PHILOX_STEP_STRIDE);\n'
        '  return 0;\n}\n')
    exe = tmp_path / "layout"
    inc = os.path.dirname(_lib.HEADER_PATH)
    subprocess.run(["gcc", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    cd, dr = _lib.ConvDesc, _lib.Dropout
    assert [int(v) for v in out[0].split()] == [ctypes.sizeof(cd), cd.src_c.offset, cd.w_stride_tap.offset,
                                                cd.w_by_src.offset, cd.accumulate.offset, cd.engine.offset,
                                                cd.tune.offset]
    assert [int(v) for v in out[1].split()] == [ctypes.sizeof(dr), dr.seed.offset, dr.stream_id.offset,
                                                dr.rate.offset, dr.step.offset, dr.mask.offset]
    assert int(out[2]) == 4096


def test_philox_streams_eager_and_graph_agree():
    """PhiloxNoise: the stream id of an eager step equals the captured-graph stream id plus
    step * M1_PHILOX_STEP_STRIDE (added on the device from the step counter), for every (pass, site)."""
    from m1b200.model.unets.engine import PhiloxNoise
    n = PhiloxNoise(seed=7, rank=1)
    seen = set()
    for step in (0, 5, 123):
        for p in PhiloxNoise.PASSES:
            for s in PhiloxNoise.SITES:
                n.step, n.step_dev = step, None
                eager = n._stream(p, s)
                n.step_dev = object()                      # capturing: the step term moves to the device
                graph = n._stream(p, s)
                assert eager == graph + step * 4096
                assert 0 <= graph < 4096
                seen.add(eager)
    assert len(seen) == 3 * len(PhiloxNoise.PASSES) * len(PhiloxNoise.SITES)        # no collisions
    assert PhiloxNoise(seed=7, rank=0).seed != n.seed                              # ranks draw different noise
