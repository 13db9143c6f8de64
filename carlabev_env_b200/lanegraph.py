"""Town01 lane graphs and the shortest-path search `rdm` / `red_light_runner` scene generation needs.

Host-side mirror of the reference's planners for this path:
  * MapGraph / GraphPlanner           -- src/planning/map_graph.py:8-95, src/planning/graph_planner.py:97-121
  * PlannerManager (which graph when) -- src/managers/scene_generator.py:18-41

The reference calls `networkx.shortest_path(G, s, t, weight="cost")`, i.e. networkx's bidirectional Dijkstra
(third-party; the container that recorded the goldens has networkx 3.6.1).  `LaneGraph.shortest_path` restates
that published algorithm on plain adjacency arrays exported in the pickles' own iteration order
(oracle/export_graphs.py), so equal-cost ties resolve the same way: alternating forward / backward expansion,
heap entries ordered by (distance, push counter), relaxation only on strict improvement, the best meeting node
updated on strict improvement, path rebuilt from the predecessor maps when a node is settled from both sides.
tests/test_host_logic.py checks it against networkx itself on random node pairs of every graph.
"""
from __future__ import annotations

import os
from heapq import heappop, heappush

import numpy as np

ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "town01_lanegraphs.npz")
RAW_TO_SURFACE = 8.0  # envs/geometry.py:6


class NoPath(Exception):
    pass


class LaneGraph:
    def __init__(self, z, key: str):
        k = key.replace("-", "_")
        self.key = key
        self.names = [str(s) for s in z[f"{k}__names"]]
        self.index = {n: i for i, n in enumerate(self.names)}
        self.pos = np.asarray(z[f"{k}__pos"], dtype=np.float64)          # raw map coordinates (x, y)
        self.pos_i32 = self.pos.astype(np.int32)                          # MapGraph.get_node_pos
        self.directed = bool(z[f"{k}__directed"])
        self._adj = []
        for nm in ("succ", "pred"):
            off, nbr, cost = z[f"{k}__{nm}_off"], z[f"{k}__{nm}_nbr"], z[f"{k}__{nm}_cost"]
            self._adj.append([list(zip(nbr[off[i]:off[i + 1]].tolist(), cost[off[i]:off[i + 1]].tolist()))
                              for i in range(len(self.names))])
        self.classes = {f[len(k) + 6:]: z[f].tolist() for f in z.files if f.startswith(f"{k}__cls_")}

    # -- MapGraph ------------------------------------------------------------------------------------------------
    def random_node(self, node_cls: str, rng) -> int:
        """MapGraph.get_random_node: rng.choice over the class list (random.Random)."""
        return rng.choice(self.classes[node_cls])

    def pos_surface(self, node: int):
        """get_node_pos_surface: int32 raw position / 8 (map_graph.py:54-58, geometry.py:17-19)."""
        return float(self.pos_i32[node, 0]) / RAW_TO_SURFACE, float(self.pos_i32[node, 1]) / RAW_TO_SURFACE

    # -- networkx.bidirectional_dijkstra, weight="cost" --------------------------------------------------------------
    def shortest_path(self, source: int, target: int) -> list[int]:
        if source == target:
            return [source]
        dists = [{}, {}]
        preds = [{source: None}, {target: None}]
        fringe = [[], []]
        seen = [{source: 0}, {target: 0}]
        c = 0
        heappush(fringe[0], (0, c, source))
        c += 1
        heappush(fringe[1], (0, c, target))
        c += 1
        finaldist, meet = None, None
        direction = 1
        while fringe[0] and fringe[1]:
            direction = 1 - direction
            dist, _, v = heappop(fringe[direction])
            dd = dists[direction]
            if v in dd:
                continue
            dd[v] = dist
            if v in dists[1 - direction]:
                fwd, cur = [], meet
                while cur is not None:
                    fwd.append(cur)
                    cur = preds[0][cur]
                fwd.reverse()
                cur = preds[1][meet]
                while cur is not None:
                    fwd.append(cur)
                    cur = preds[1][cur]
                return fwd
            sd, so = seen[direction], seen[1 - direction]
            for w, cost in self._adj[direction][v]:
                length = dist + cost
                if w in dd:
                    if length < dd[w]:
                        raise ValueError("Contradictory paths found: negative weights?")
                elif w not in sd or length < sd[w]:
                    sd[w] = length
                    heappush(fringe[direction], (length, c, w))
                    c += 1
                    preds[direction][w] = v
                    if w in so:
                        total = length + so[w]
                        if finaldist is None or finaldist > total:
                            finaldist, meet = total, w
        raise NoPath(f"No path between {self.names[source]} and {self.names[target]}.")

    def find_path(self, start: int, end: int, threshold: float = 10.0) -> list[int]:
        """GraphPlanner.find_path (graph_planner.py:97-121): shortest path thinned so that consecutive kept
        nodes are more than `threshold` raw pixels apart; [] when there is no path."""
        try:
            path = self.shortest_path(start, end)
        except NoPath:
            return []
        merged, last = [], None
        for n in path:
            p = self.pos[n]
            if last is None or float(np.linalg.norm(p - last)) > threshold:
                merged.append(n)
                last = p
        return merged


_GRAPHS: dict = {}


def load_graph(key: str) -> LaneGraph:
    """key in {"vehicle-full", "vehicle", "vehicle-L", "vehicle-R"} (PlannerManager.graphs)."""
    if key not in _GRAPHS:
        with np.load(ASSET) as z:
            _GRAPHS[key] = LaneGraph(z, key)
    return _GRAPHS[key]
