"""The reference's own test file, call for call: /root/reference/test/runtests.jl:4-36 (nine `@test
solveWithTrustedFunctions(...)`, same arguments, same keyword names), then the `@assert`s and printed verdicts of
its example scripts.  Everything goes through the public mirror `ecneproject_b200.solveWithTrustedFunctions`, i.e.
native reader -> native abstraction -> ecne_solve on the GPU."""
import pytest

from ecneproject_b200 import fixtures, solveWithTrustedFunctions

pytestmark = pytest.mark.gpu

F = fixtures.path
PED = [F("tornadocash_circuits/Pedersen248@pedersen.r1cs"), F("tornadocash_circuits/Pedersen496@pedersen.r1cs")]


def test_unused_argument():  # runtests.jl:4-6
    assert solveWithTrustedFunctions(F("straightforward.r1cs"), "Unused Argument", printRes=False)


def test_trivial_multiplication():  # runtests.jl:8-10
    assert solveWithTrustedFunctions(F("trivial_mult.r1cs"), "*", printRes=False)


def test_big_mult():  # runtests.jl:12-14
    assert solveWithTrustedFunctions(F("bigmult86_3.r1cs"), "bigmult(86,3)", printRes=False)


def test_poseidon():  # runtests.jl:16-18
    assert solveWithTrustedFunctions(F("poseidon.r1cs"), "poseidon", printRes=False)


def test_3x3_multiplexer():  # runtests.jl:20-22
    assert solveWithTrustedFunctions(F("multiplexer_33.r1cs"), "multiplexer(3,3)", printRes=False)


def test_tornadocash_circuits():  # runtests.jl:24-27
    assert solveWithTrustedFunctions(F("tornadocash_circuits/commitHasher.r1cs"), "CommitmentHasher",
                                     trusted_r1cs=PED, trusted_r1cs_names=["Pedersen248", "Pedersen496"], printRes=False)
    assert solveWithTrustedFunctions(F("tornadocash_circuits/merkleTree.r1cs"), "MerkleTreeChecker", printRes=False)


def test_tornadocash_withdraw_circuits():  # runtests.jl:29-31
    assert solveWithTrustedFunctions(F("tornadocash_circuits/withdraw.r1cs"), "Withdraw",
                                     trusted_r1cs=PED, trusted_r1cs_names=["Pedersen248", "Pedersen496"], printRes=False)


def test_secp_add_unequal_given_bigmultmodp_biglessthan():  # runtests.jl:34-36
    assert solveWithTrustedFunctions(F("secp256k1.r1cs"), "secpAddUnequal",
                                     trusted_r1cs=[F("bigmultmodp.r1cs"), F("biglessthan.r1cs")],
                                     trusted_r1cs_names=["BigMultModP", "BigLessThan"], secp_solve=True, printRes=False)


def test_example_ecdsa_secp_abstraction():  # examples/ecdsa_secp_abstraction.jl:4 (@assert ... == true)
    assert solveWithTrustedFunctions(F("ecdsa.r1cs"), "ECDSAPrivToPub", trusted_r1cs=[F("secp256k1.r1cs")],
                                     trusted_r1cs_names=["Secp256k1AddUnequal"], printRes=False) is True


def test_example_division_and_boundcheck(capsys):  # README.md:106, examples/division.jl, examples/boundcheck.jl
    assert solveWithTrustedFunctions(F("target/division.r1cs"), "division") is False
    assert "R1CS function division has potentially unsound constraints" in capsys.readouterr().out
    assert solveWithTrustedFunctions(F("bad_bd_check.r1cs"), "bad_bd_check", printRes=False) is False
    assert solveWithTrustedFunctions(F("good_bd_check.r1cs"), "good_bd_check") is True
    assert "R1CS function good_bd_check has sound constraints (No trusted functions needed!)" in capsys.readouterr().out
