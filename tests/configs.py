"""Run configurations = (main circuit, trusted circuits, flags), named as in SURVEY.md Appendix B.

`pinned` is the Boolean the reference itself asserts for the config, with the asserting file:line.
"""
from ecneproject_b200 import fixtures

CONFIGS = {}

for _n in fixtures.circomlib():
    CONFIGS["circomlib/" + _n] = {"main": f"ecne_circomlib_tests/{_n}.r1cs"}

_PED = ["tornadocash_circuits/Pedersen248@pedersen.r1cs", "tornadocash_circuits/Pedersen496@pedersen.r1cs"]
_PEDN = ["Pedersen248", "Pedersen496"]

CONFIGS.update({
    "root/straightforward": {"main": "straightforward.r1cs", "pinned": (True, "test/runtests.jl:5")},
    "root/trivial_mult": {"main": "trivial_mult.r1cs", "pinned": (True, "test/runtests.jl:9")},
    "root/bigmult86_3": {"main": "bigmult86_3.r1cs", "pinned": (True, "test/runtests.jl:13")},
    "root/poseidon": {"main": "poseidon.r1cs", "pinned": (True, "test/runtests.jl:17")},
    "root/multiplexer_33": {"main": "multiplexer_33.r1cs", "pinned": (True, "test/runtests.jl:21")},
    "tornado/commitHasher+pedersen": {"main": "tornadocash_circuits/commitHasher.r1cs", "trusted": _PED,
                                      "trusted_names": _PEDN, "pinned": (True, "test/runtests.jl:25")},
    "tornado/merkleTree": {"main": "tornadocash_circuits/merkleTree.r1cs", "pinned": (True, "test/runtests.jl:26")},
    "tornado/withdraw+pedersen": {"main": "tornadocash_circuits/withdraw.r1cs", "trusted": _PED,
                                  "trusted_names": _PEDN, "pinned": (True, "test/runtests.jl:30")},
    "secp256k1+bmmp+blt": {"main": "secp256k1.r1cs", "trusted": ["bigmultmodp.r1cs", "biglessthan.r1cs"],
                           "trusted_names": ["BigMultModP", "BigLessThan"], "secp_solve": True,
                           "pinned": (True, "test/runtests.jl:35")},
    "target/division": {"main": "target/division.r1cs", "pinned": (False, "README.md:106")},
    "root/good_bd_check": {"main": "good_bd_check.r1cs", "pinned": (True, "examples/boundcheck.jl (name)")},
    "root/bad_bd_check": {"main": "bad_bd_check.r1cs", "pinned": (False, "examples/boundcheck.jl (name)")},
    # unpinned extras
    "tornado/withdraw": {"main": "tornadocash_circuits/withdraw.r1cs"},
    "tornado/commitHasher": {"main": "tornadocash_circuits/commitHasher.r1cs"},
    "root/secp256k1": {"main": "secp256k1.r1cs"},
    "root/bigmultmodp": {"main": "bigmultmodp.r1cs"},
    "root/biglessthan": {"main": "biglessthan.r1cs"},
    "root/bigmultshortlong86_3": {"main": "bigmultshortlong86_3.r1cs"},
    "tornado/Pedersen248": {"main": "tornadocash_circuits/Pedersen248@pedersen.r1cs"},
    "tornado/Pedersen496": {"main": "tornadocash_circuits/Pedersen496@pedersen.r1cs"},
    "benchmarks/bigmod_5_2": {"main": "Circom_Functions/benchmarks/bigmod_5_2.r1cs"},
    "benchmarks/bigmod_10_2": {"main": "Circom_Functions/benchmarks/bigmod_10_2.r1cs"},
    "benchmarks/bigmod_86_3": {"main": "Circom_Functions/benchmarks/bigmod_86_3.r1cs"},
    # the headline workload (BASELINE.json configs[4]); big = minutes of CPU oracle time
    "ecdsa+secp256k1": {"main": "ecdsa.r1cs", "trusted": ["secp256k1.r1cs"], "trusted_names": ["Secp256k1AddUnequal"],
                        "big": True, "pinned": (True, "examples/ecdsa_secp_abstraction.jl:4")},
    "ecdsa": {"main": "ecdsa.r1cs", "big": True},
})
