"""CPU: the step programs of the warp-cooperative pairing (tools/vmgen.py -> harmony_b200/csrc/vm_programs.cuh), run by the
generator's reference interpreter, against the oracle (oracle/pyref.py): valid and invalid signature checks, multi-signer aggregate,
structural limits of the encoding, and that the committed header is what the generator produces."""
import importlib.util, os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

@pytest.fixture(scope="module")
def vm():
    spec = importlib.util.spec_from_file_location("vmgen", os.path.join(ROOT, "tools", "vmgen.py")); m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m

@pytest.fixture(scope="module")
def progs(vm): return vm.compile_all()

def _instance(sk, msg):
    import pyref as o
    pk = o.get_public_key(sk); sig = o.sign_hash(sk, msg); hm = o.map_to_g2(msg)
    gen = o.pt_affine(o.FP, o.G1_GEN); npk = o.pt_affine(o.FP, o.pt_neg(o.FP, pk))
    sa = o.pt_affine(o.FP2, sig); ha = o.pt_affine(o.FP2, hm)
    return gen, sa, npk, ha

def test_vm_pairing_valid_and_invalid(vm, progs):
    import pyref as o
    gen, sa, npk, ha = _instance(0x1234567890abcdef1234567890abcdef, b"vm-test-message-0123456789abcdef")
    assert vm.vm_pairing_is_one(progs, gen, sa, npk, ha) is True
    assert o.pairing_product_is_one([(o.pt_from_affine(o.FP, *gen), o.pt_from_affine(o.FP2, *sa)), (o.pt_from_affine(o.FP, *npk), o.pt_from_affine(o.FP2, *ha))])
    # wrong message / wrong key / swapped operands
    _, _, _, hb = _instance(0x1234567890abcdef1234567890abcdef, b"another message.................")
    assert vm.vm_pairing_is_one(progs, gen, sa, npk, hb) is False
    _, _, npk2, _ = _instance(0x77, b"vm-test-message-0123456789abcdef")
    assert vm.vm_pairing_is_one(progs, gen, sa, npk2, ha) is False

def test_vm_final_exp_matches_oracle_value(vm, progs):
    """The Fp12 value after the VM's final exponentiation equals the oracle's f^(3 (p^12 - 1) / r) on a Miller-loop output
    (same cube-of-the-pairing convention as pairing.cuh), checked through  r_vm == 1  <=>  r_oracle == 1 and by cubing."""
    import pyref as o
    gen, sa, npk, ha = _instance(0xabcdef, b"fe-check........................")
    s = vm.fresh_slots()
    for reg, v in (("P1X", (gen[0], 0)), ("P1Y", (gen[1], 0)), ("Q1X", sa[0]), ("Q1Y", sa[1]), ("P2X", (npk[0], 0)), ("P2Y", (npk[1], 0)), ("Q2X", ha[0]), ("Q2Y", ha[1])):
        s[vm.SLOT[reg]] = v
    vm.run_program(progs["ML_INIT"], s)
    for i in range(62, -1, -1):
        vm.run_program(progs["ML_DBL"], s)
        if (vm.Z_ABS >> i) & 1: vm.run_program(progs["ML_ADD"], s)
    # tower order (c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2) -> the oracle's w-power order (w^0 .. w^5)
    f_t = [s[vm.SLOT[r]] for r in vm.r6("F")]
    f_w = tuple(f_t[vm.SLOT_OF_W[k]] for k in range(6))
    vm.final_exp_vm(progs, s)
    r_t = [s[vm.SLOT[r]] for r in vm.r6("ACC")]
    r_w = tuple(r_t[vm.SLOT_OF_W[k]] for k in range(6))
    want = o.f12_pow(f_w, 3 * ((o.P ** 12 - 1) // o.R))
    assert r_w == want == o.F12_ONE

def test_vm_encoding_limits(vm, progs):
    st = vm.stats(progs)
    assert st["ML_DBL"]["mul_ops"] == 66 and st["CYCSQR"]["sqr_ops"] == 9 and st["MULX"]["mul_ops"] == 18
    for name, p in progs.items():
        for cls, ins in p.steps:
            assert 1 <= len(ins) <= vm.NPAIR
            for i in ins:
                assert 0 <= i[0] < vm.NSLOTS < 255
                if cls == vm.OP_LIN:
                    assert 1 <= len(i[1]) <= vm.MAXT
                    for sl, M in i[1]: assert 0 <= sl < vm.NSLOTS and all(abs(c) <= vm.CMAX for c in M)
                else: assert 0 <= i[1] < vm.NSLOTS and 0 <= i[2] < vm.NSLOTS

def test_vm_header_is_current(vm, progs, tmp_path):
    out = tmp_path / "vm_programs.cuh"
    vm.emit(progs, str(out))
    assert out.read_text() == open(os.path.join(ROOT, "harmony_b200", "csrc", "vm_programs.cuh")).read(), "run python tools/vmgen.py"
