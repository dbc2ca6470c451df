"""Roadmap CSV format of the reference: one line per node,
    (q1,...,qD):(n1...),(n2...);
reader semantics of utils/pb_data_utils.py:18-52 (lines without ':' -- isolated nodes --
are skipped; a missing file gives an empty roadmap), writer format of
res/images/vis_taskspace_sampling_3D.py:479-500."""
import os
import re

_FLOAT = re.compile(r'-?\d+(?:\.\d+)?(?:[eE]-?\d+)?')


def _tuple(text):
    return tuple(float(x) for x in _FLOAT.findall(text))


def load_roadmap(filename):
    roadmap = {}
    if not os.path.exists(filename):
        print(f"File not found: {filename}")
        return roadmap
    with open(filename, 'r') as f:
        for line in f:
            parts = line.strip().split(":")
            if len(parts) != 2:
                continue
            roadmap[_tuple(parts[0])] = [_tuple(nb) for nb in parts[1].split('),(')]
    return roadmap


def save_roadmap(roadmap, filename):
    """roadmap: {q tuple: [neighbour q tuples]}"""
    with open(filename, 'w') as f:
        for node, neighbors in roadmap.items():
            line = "(" + ",".join(str(v) for v in node) + "):"
            line += ",".join("(" + ",".join(str(v) for v in nb) + ")" for nb in neighbors)
            f.write(line + ";\n")


def planner_roadmap(prm, r_index=0):
    """{q: [neighbour qs]} of one robot of a DecoupledPRM-style planner."""
    d = prm.node_id_q_dicts[r_index]
    return {d[i]: [d[j] for j in nbrs] for i, nbrs in prm.edge_dicts[r_index].items() if i >= 0}
