"""Classifier dependency graph (host logic; mirrors ``allophant/attribute_graph.py:17-199``).

Decides the order in which the classifier heads are evaluated and therefore the
order of ``Predictions.outputs``.  ``sort()`` yields nodes in the same order as the
reference's iterative Tarjan walk: a depth-first post-order that starts at node 0,
follows dependencies in their listed order and restarts at the lowest unvisited
index.  Cycles raise ``DependencyCycleError`` naming every node of the cycle.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, Iterator, List, Mapping, Optional

from .config import MultiheadAttentionConfig, ProjectionEntryConfig


class DependencyCycleError(Exception):
    """Raised when a dependency cycle is detected"""


@dataclass
class AttributeNode:
    name: str
    size: int
    time_layer_config: Optional[MultiheadAttentionConfig] = None
    dependencies: List[str] = field(default_factory=list)

    def with_offset(self, offset: int = 1) -> "AttributeNode":
        return AttributeNode(self.name, self.size + offset, self.time_layer_config, self.dependencies)


class AttributeGraph:
    _nodes: List[AttributeNode]
    _node_indices: Dict[str, int]
    _edges: List[List[int]]

    def __init__(self, nodes: Iterable[AttributeNode]) -> None:
        self._nodes = []
        self._node_indices = {}
        for index, node in enumerate(nodes):
            self._nodes.append(node)
            self._node_indices[node.name] = index
        self._edges = [
            [
                self._node_indices[dependency]
                for dependency in node.dependencies
                if not ProjectionEntryConfig.OUTPUT_PATTERN.match(dependency)
            ]
            for node in self._nodes
        ]

    # -- (de)serialisation of the checkpoint entry `attribute_graph` (attribute_graph.py:202-242) --
    @classmethod
    def from_state(cls, state: Mapping[str, Any]) -> "AttributeGraph":
        graph = cls.__new__(cls)
        graph._nodes = []
        for node in state["nodes"]:
            time_layer = node.get("time_layer_config")
            graph._nodes.append(
                AttributeNode(
                    node["name"],
                    int(node["size"]),
                    None
                    if time_layer is None
                    else MultiheadAttentionConfig(
                        time_layer.get("num_heads", 1), time_layer.get("positional_embeddings", False)
                    ),
                    list(node.get("dependencies", [])),
                )
            )
        graph._node_indices = {name: int(index) for name, index in state["node_indices"].items()}
        graph._edges = [list(map(int, edges)) for edges in state["edges"]]
        return graph

    def state(self) -> Dict[str, Any]:
        return {
            "nodes": [
                {
                    "name": node.name,
                    "size": node.size,
                    "time_layer_config": None if node.time_layer_config is None else node.time_layer_config.dump(),
                    "dependencies": list(node.dependencies),
                }
                for node in self._nodes
            ],
            "node_indices": dict(self._node_indices),
            "edges": [list(edges) for edges in self._edges],
        }

    def sizes(self) -> Iterator[int]:
        return (node.size for node in self._nodes)

    def names(self) -> Iterator[str]:
        return (node.name for node in self._nodes)

    @property
    def nodes(self) -> List[AttributeNode]:
        return self._nodes

    def get(self, node: "str | int") -> Optional[AttributeNode]:
        if isinstance(node, str):
            node_index = self._node_indices.get(node)
            if node_index is None:
                return None
            node = node_index
        return self._nodes[node]

    def __getitem__(self, node: "str | int") -> AttributeNode:
        if isinstance(node, str):
            node = self._node_indices[node]
        return self._nodes[node]

    def __iter__(self) -> Iterator[AttributeNode]:
        return iter(self._nodes)

    def __contains__(self, node_name: str) -> bool:
        return node_name in self._node_indices

    def __len__(self) -> int:
        return len(self._nodes)

    def sort(self) -> Iterator[AttributeNode]:
        """Nodes in reverse topological order (dependencies before dependents)."""
        WHITE, GREY, BLACK = 0, 1, 2
        colour = [WHITE] * len(self._nodes)
        order: List[int] = []
        for root in range(len(self._nodes)):
            if colour[root] != WHITE:
                continue
            # explicit stack of (node, next edge position); `path` is the current DFS chain
            stack = [(root, 0)]
            colour[root] = GREY
            path = [root]
            while stack:
                node, position = stack.pop()
                if position < len(self._edges[node]):
                    stack.append((node, position + 1))
                    target = self._edges[node][position]
                    if colour[target] == GREY:
                        cycle = path[path.index(target) :]
                        raise DependencyCycleError(
                            "Dependency cycle detected: " + " -> ".join(self._nodes[i].name for i in reversed(cycle))
                        )
                    if colour[target] == WHITE:
                        colour[target] = GREY
                        path.append(target)
                        stack.append((target, 0))
                else:
                    colour[node] = BLACK
                    path.pop()
                    order.append(node)
        return (self._nodes[index] for index in order)
