"""CPU: tensor-network index maps bit-exact against the reference's gen_tensor_networks output
(fixtures: tests/golden/tn_index_maps.json, made by tests/golden/generate_golden.py)."""
import pytest

import tedq_b200 as qb
from conftest import load_golden
from tedq_b200 import tn_index
from tedq_b200 import workloads as W

MAPS = load_golden("tn_index_maps.json")


@pytest.mark.parametrize("rec", MAPS, ids=lambda r: r["spec"]["name"])
def test_index_maps_bit_exact(rec):
    circ = W.build_circuit(rec["spec"], qb)
    nets = tn_index.networks_of_circuit(circ)
    assert len(nets) == len(rec["networks"])
    for net, ref in zip(nets, rec["networks"]):
        ins, out = net.symbols()
        assert ins == ref["inputs"]
        assert out == ref["output"]
        assert net.size_keys() == ref["size_keys"]


def test_survey_worked_examples():
    """SURVEY.md 8a worked examples (produced by the reference itself)."""
    nets = tn_index.index_maps(2, [[0], [1]], [("expval", [[0]]), ("state", None)])
    assert nets[0].symbols() == ([['a'], ['b'], ['c', 'a'], ['d', 'b'], ['e', 'c'], ['f', 'd'], ['g', 'e'], ['g'], ['f']], [])
    assert nets[1].symbols() == ([['a'], ['b'], ['c', 'a'], ['d', 'b']], ['c', 'd'])
    nets = tn_index.index_maps(2, [[0], [0, 1], [1]], [("expval", [[1]])])
    assert nets[0].symbols()[0] == [['a'], ['b'], ['c', 'a'], ['d', 'e', 'c', 'b'], ['f', 'e'], ['g', 'f'], ['h', 'g'],
                                    ['i', 'j', 'd', 'h'], ['k', 'i'], ['k'], ['j']]
    assert tn_index.symbol(1) == 'b' and tn_index.symbol(200) == chr(340)
