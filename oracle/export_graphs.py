"""Convert the reference's Town01 lane graphs (networkx pickles, DATA assets) into one plain-array file.

TEST / DATA INFRASTRUCTURE, run in the build container only:  python oracle/export_graphs.py
Writes carlabev_env_b200/assets/town01_lanegraphs.npz (read by carlabev_env_b200/lanegraph.py).

What is kept, per planner of PlannerManager (managers/scene_generator.py:18-41) that `rdm` and
`red_light_runner` scene generation touch:
  * node names in the pickle's insertion order (ints are tagged "#<n>"), raw positions as float64,
  * the sampling lists MapGraph.get_lane_nodes builds (planning/map_graph.py:22-45), taken from the
    reference's own loader so that `rng.choice(list)` indexes the same nodes,
  * adjacency in the pickle's iteration order (successors and predecessors; the same table twice for the
    undirected graph) with `cost` (absent -> 1, networkx's default for weight="cost").
Iteration order is what fixes the tie-breaks of the shortest-path search, so it is exported verbatim.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_loader import REFERENCE_ROOT, load_reference  # noqa: E402

OUT = os.path.join(ROOT, "carlabev_env_b200", "assets", "town01_lanegraphs.npz")
GRAPHS = {  # key in PlannerManager.graphs -> pickle
    "vehicle-full": "town01-vehicles-100.pkl",
    "vehicle": "town01-vehicles-2lanes-100.pkl",
    "vehicle-L": "town01-vehicles-left-100.pkl",
    "vehicle-R": "town01-vehicles-right-100.pkl",
}


def tag(n):
    return f"#{n}" if isinstance(n, (int, np.integer)) else str(n)


def csr(adj, index):
    off, nbr, cost = [0], [], []
    for n in index:
        for w, d in adj[n].items():
            nbr.append(index[w])
            cost.append(float(d.get("cost", 1)))
        off.append(len(nbr))
    return np.array(off, np.int32), np.array(nbr, np.int32), np.array(cost, np.float64)


def main():
    load_reference()
    from CarlaBEV.src.planning.graph_planner import GraphPlanner

    out = {}
    for key, fname in GRAPHS.items():
        gp = GraphPlanner(os.path.join(REFERENCE_ROOT, "CarlaBEV", "assets", "Town01", fname))
        G = gp.G
        index = {n: i for i, n in enumerate(G.nodes)}
        k = key.replace("-", "_")
        out[f"{k}__names"] = np.array([tag(n) for n in G.nodes])
        out[f"{k}__pos"] = np.array([np.asarray(G.nodes[n]["pos"], dtype=np.float64) for n in G.nodes])
        # get_node_pos casts to int32 (map_graph.py:54-55); red_light_running.py reads the float position
        out[f"{k}__directed"] = np.bool_(G.is_directed())
        succ = G._succ if G.is_directed() else G._adj
        pred = G._pred if G.is_directed() else G._adj
        for nm, adj in (("succ", succ), ("pred", pred)):
            off, nbr, cost = csr(adj, index)
            out[f"{k}__{nm}_off"], out[f"{k}__{nm}_nbr"], out[f"{k}__{nm}_cost"] = off, nbr, cost
        for cls, nodes in gp.nodes.items():
            out[f"{k}__cls_{cls}"] = np.array([index[n] for n in nodes], dtype=np.int32)
        print(key, G.number_of_nodes(), "nodes", G.number_of_edges(), "edges",
              {c: len(v) for c, v in gp.nodes.items() if len(v)})
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, f"{os.path.getsize(OUT) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
