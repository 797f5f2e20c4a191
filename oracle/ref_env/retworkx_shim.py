"""A small pure-Python stand-in for ``retworkx`` (TEST INFRASTRUCTURE ONLY).

Qiskit Terra 0.18's transpiler stores its DAG in ``retworkx.PyDAG`` (qiskit/dagcircuit/dagcircuit.py:30,77);
retworkx is a compiled Rust extension that is not installed here and cannot be installed (no network),
which is why ``transpile()`` / ``execute()`` cannot run in this container.  This
module implements exactly the part of the retworkx API that ``DAGCircuit`` and the level-0/1 pass
managers use for a backend without a coupling map (which is what BasicAer-like simulators are), so that
the reference's own ``execute(circuit, backend)`` can be driven end to end in the CPU test-suite against
the drop-in backends.  It is registered as ``retworkx`` by ``ref_import.import_reference()``.

It is not part of the product, not an optimised graph library, and only as faithful as the call sites
listed below need it to be (indices are stable integers, parallel edges are allowed, node payloads are
arbitrary Python objects).
"""
import heapq
import itertools


class DAGHasCycle(Exception):
    pass


class NoSuitableNeighbors(Exception):
    pass


class NullGraph(Exception):
    pass


class NoEdgeBetweenNodes(Exception):
    pass


class InvalidNode(Exception):
    pass


class PyDiGraph:
    """Directed multigraph with stable integer node indices (dagcircuit.py:77 and call sites)."""

    def __init__(self, check_cycle=False, multigraph=True):
        self.check_cycle = check_cycle
        self.multigraph = multigraph
        self._nodes = {}      # index -> payload
        self._next_node = 0
        self._edges = {}      # edge id -> (u, v, payload)
        self._next_edge = 0
        self._out = {}        # u -> list of edge ids (insertion order)
        self._in = {}         # v -> list of edge ids

    # -- nodes ---------------------------------------------------------------------------------
    def add_node(self, obj):
        idx = self._next_node
        self._next_node += 1
        self._nodes[idx] = obj
        self._out[idx] = []
        self._in[idx] = []
        return idx

    def add_nodes_from(self, objs):
        return [self.add_node(o) for o in objs]

    def remove_node(self, idx):
        if idx not in self._nodes:
            return
        for eid in list(self._out[idx]) + list(self._in[idx]):
            self._drop_edge(eid)
        del self._nodes[idx], self._out[idx], self._in[idx]

    def nodes(self):
        return [self._nodes[i] for i in sorted(self._nodes)]

    def node_indexes(self):
        return sorted(self._nodes)

    node_indices = node_indexes

    def __getitem__(self, idx):
        try:
            return self._nodes[idx]
        except KeyError:
            raise IndexError("No node found for index") from None

    def __setitem__(self, idx, value):
        if idx not in self._nodes:
            raise IndexError("No node found for index")
        self._nodes[idx] = value

    def __len__(self):
        return len(self._nodes)

    def get_node_data(self, idx):
        return self[idx]

    # -- edges ---------------------------------------------------------------------------------
    def add_edge(self, u, v, obj):
        if u not in self._nodes or v not in self._nodes:
            raise InvalidNode("One of the endpoints of the edge does not exist in graph")
        eid = self._next_edge
        self._next_edge += 1
        self._edges[eid] = (u, v, obj)
        self._out[u].append(eid)
        self._in[v].append(eid)
        return eid

    def add_edges_from(self, edges):
        return [self.add_edge(u, v, w) for u, v, w in edges]

    def add_edges_from_no_data(self, edges):
        return [self.add_edge(u, v, None) for u, v in edges]

    def _drop_edge(self, eid):
        u, v, _ = self._edges.pop(eid)
        self._out[u].remove(eid)
        self._in[v].remove(eid)

    def remove_edge(self, u, v):
        u = self._index_of(u)
        v = self._index_of(v)
        for eid in self._out.get(u, []):
            if self._edges[eid][1] == v:
                self._drop_edge(eid)
                return
        raise NoEdgeBetweenNodes("No edge found between nodes")

    def remove_edge_from_index(self, eid):
        if eid in self._edges:
            self._drop_edge(eid)

    def _index_of(self, node):
        # DAGCircuit sometimes passes the node object instead of its id (dagcircuit.py:1082)
        if isinstance(node, int):
            return node
        return getattr(node, "_node_id", node)

    def edges(self):
        return [self._edges[e][2] for e in sorted(self._edges)]

    def edge_list(self):
        return [(self._edges[e][0], self._edges[e][1]) for e in sorted(self._edges)]

    def weighted_edge_list(self):
        return [self._edges[e] for e in sorted(self._edges)]

    def has_edge(self, u, v):
        return any(self._edges[e][1] == v for e in self._out.get(u, []))

    def get_edge_data(self, u, v):
        for eid in self._out.get(u, []):
            if self._edges[eid][1] == v:
                return self._edges[eid][2]
        raise NoEdgeBetweenNodes("No edge found between nodes")

    def get_all_edge_data(self, u, v):
        out = [self._edges[e][2] for e in self._out.get(u, []) if self._edges[e][1] == v]
        if not out:
            raise NoEdgeBetweenNodes("No edge found between nodes")
        return out

    def in_edges(self, idx):
        return [self._edges[e] for e in reversed(self._in.get(self._index_of(idx), []))]

    def out_edges(self, idx):
        return [self._edges[e] for e in reversed(self._out.get(self._index_of(idx), []))]

    def in_degree(self, idx):
        return len(self._in.get(idx, []))

    def out_degree(self, idx):
        return len(self._out.get(idx, []))

    # -- neighbourhood ---------------------------------------------------------------------------
    def successor_indices(self, idx):
        seen, out = set(), []
        for eid in reversed(self._out.get(self._index_of(idx), [])):
            v = self._edges[eid][1]
            if v not in seen:
                seen.add(v)
                out.append(v)
        return out

    def predecessor_indices(self, idx):
        seen, out = set(), []
        for eid in reversed(self._in.get(self._index_of(idx), [])):
            u = self._edges[eid][0]
            if u not in seen:
                seen.add(u)
                out.append(u)
        return out

    def successors(self, idx):
        return [self._nodes[v] for v in self.successor_indices(idx)]

    def predecessors(self, idx):
        return [self._nodes[u] for u in self.predecessor_indices(idx)]

    def neighbors(self, idx):
        return self.successor_indices(idx)

    def find_adjacent_node_by_edge(self, idx, predicate):
        for eid in reversed(self._out.get(idx, [])):
            _, v, w = self._edges[eid]
            if predicate(w):
                return self._nodes[v]
        raise NoSuitableNeighbors("No suitable neighbor")

    def find_predecessors_by_edge(self, idx, predicate):
        out = []
        for eid in reversed(self._in.get(idx, [])):
            u, _, w = self._edges[eid]
            if predicate(w):
                out.append(self._nodes[u])
        return out

    def find_successors_by_edge(self, idx, predicate):
        out = []
        for eid in reversed(self._out.get(idx, [])):
            _, v, w = self._edges[eid]
            if predicate(w):
                out.append(self._nodes[v])
        return out

    # -- rewiring used by DAGCircuit -------------------------------------------------------------
    def insert_node_on_in_edges_multiple(self, node, ref_nodes):
        """dagcircuit.py:430 (apply_operation_back): every edge (p -> ref, w) becomes p -> node -> ref."""
        for ref in ref_nodes:
            for eid in list(self._in[ref]):
                p, _, w = self._edges[eid]
                self._drop_edge(eid)
                self.add_edge(p, node, w)
                self.add_edge(node, ref, w)

    def insert_node_on_out_edges_multiple(self, node, ref_nodes):
        """dagcircuit.py:468 (apply_operation_front): every edge (ref -> s, w) becomes ref -> node -> s."""
        for ref in ref_nodes:
            for eid in list(self._out[ref]):
                _, s, w = self._edges[eid]
                self._drop_edge(eid)
                self.add_edge(ref, node, w)
                self.add_edge(node, s, w)

    def insert_node_on_in_edges(self, node, ref):
        self.insert_node_on_in_edges_multiple(node, [ref])

    def insert_node_on_out_edges(self, node, ref):
        self.insert_node_on_out_edges_multiple(node, [ref])

    def remove_node_retain_edges(self, idx, use_outgoing=None, condition=None):
        """dagcircuit.py:1322 (remove_op_node): connect each predecessor to each successor whose edge
        payloads satisfy ``condition``, then drop the node."""
        ins = [self._edges[e] for e in self._in[idx]]
        outs = [self._edges[e] for e in self._out[idx]]
        for p, _, w_in in ins:
            for _, s, w_out in outs:
                if condition is None or condition(w_in, w_out):
                    self.add_edge(p, s, w_out if use_outgoing else w_in)
        self.remove_node(idx)

    def copy(self):
        new = type(self)(self.check_cycle, self.multigraph)
        new._nodes = dict(self._nodes)
        new._next_node = self._next_node
        new._edges = dict(self._edges)
        new._next_edge = self._next_edge
        new._out = {k: list(v) for k, v in self._out.items()}
        new._in = {k: list(v) for k, v in self._in.items()}
        return new

    def to_dot(self, *args, **kwargs):  # drawing is out of scope
        raise NotImplementedError("retworkx shim: to_dot is not available")


PyDAG = PyDiGraph


class PyGraph(PyDiGraph):
    """Undirected variant: an edge is stored once and seen from both ends (opflow/converters/abelian_grouper.py:108)."""

    def neighbors(self, idx):
        seen, out = set(), []
        for eid in self._out.get(idx, []) + self._in.get(idx, []):
            u, v, _ = self._edges[eid]
            other = v if u == idx else u
            if other not in seen:
                seen.add(other)
                out.append(other)
        return out

    def degree(self, idx):
        return len(self._out.get(idx, [])) + len(self._in.get(idx, []))


def graph_greedy_color(graph):
    """Greedy colouring, largest degree first (abelian_grouper.py:112); returns {node index: colour}."""
    colors = {}
    for node in sorted(graph._nodes, key=lambda n: (-graph.degree(n), n)):
        used = {colors[m] for m in graph.neighbors(node) if m in colors}
        color = 0
        while color in used:
            color += 1
        colors[node] = color
    return colors


# -- algorithms ------------------------------------------------------------------------------------
def _in_degrees(graph):
    return {n: len(graph._in[n]) for n in graph._nodes}


def topological_sort(graph):
    """Node indices in a topological order (dagcircuit.py:105)."""
    indeg = _in_degrees(graph)
    ready = [n for n in sorted(indeg) if indeg[n] == 0]
    out = []
    while ready:
        n = ready.pop()
        out.append(n)
        for eid in graph._out[n]:
            v = graph._edges[eid][1]
            indeg[v] -= 1
            if indeg[v] == 0:
                ready.append(v)
    if len(out) != len(graph._nodes):
        raise DAGHasCycle("Sort encountered a cycle")
    return out


def lexicographical_topological_sort(graph, key):
    """Node payloads in topological order, ties broken by ``key(payload)`` (dagcircuit.py:955)."""
    indeg = _in_degrees(graph)
    counter = itertools.count()
    heap = [(key(graph._nodes[n]), next(counter), n) for n in sorted(indeg) if indeg[n] == 0]
    heapq.heapify(heap)
    out = []
    while heap:
        _, _, n = heapq.heappop(heap)
        out.append(graph._nodes[n])
        for eid in graph._out[n]:
            v = graph._edges[eid][1]
            indeg[v] -= 1
            if indeg[v] == 0:
                heapq.heappush(heap, (key(graph._nodes[v]), next(counter), v))
    if len(out) != len(graph._nodes):
        raise DAGHasCycle("Sort encountered a cycle")
    return out


def layers(graph, first_layer):
    """Greedy layers of node payloads starting from ``first_layer`` indices (dagcircuit.py:1454)."""
    out = []
    indeg = _in_degrees(graph)
    cur = list(first_layer)
    seen = set(cur)
    while cur:
        out.append([graph._nodes[n] for n in cur])
        nxt = []
        for n in cur:
            for eid in graph._out[n]:
                v = graph._edges[eid][1]
                indeg[v] -= 1
                if indeg[v] == 0 and v not in seen:
                    seen.add(v)
                    nxt.append(v)
        cur = nxt
    return out


def collect_runs(graph, filter_fn):
    """Maximal chains of nodes accepted by ``filter_fn`` where each node has exactly one distinct successor
    (dagcircuit.py:1472,1489: runs of 1-qubit gates)."""
    out = []
    seen = set()
    for n in topological_sort(graph):
        node = graph._nodes[n]
        if n in seen or not filter_fn(node):
            continue
        seen.add(n)
        group = [node]
        succ = graph.successor_indices(n)
        while len(succ) == 1 and succ[0] not in seen and filter_fn(graph._nodes[succ[0]]):
            seen.add(succ[0])
            group.append(graph._nodes[succ[0]])
            succ = graph.successor_indices(succ[0])
        out.append(group)
    return out


def _reach(graph, start, forward):
    seen = set()
    stack = [start]
    while stack:
        n = stack.pop()
        eids = graph._out[n] if forward else graph._in[n]
        for eid in eids:
            v = graph._edges[eid][1 if forward else 0]
            if v not in seen:
                seen.add(v)
                stack.append(v)
    return seen


def ancestors(graph, node):
    return _reach(graph, node, False)


def descendants(graph, node):
    return _reach(graph, node, True)


def bfs_successors(graph, node):
    out = []
    seen = {node}
    queue = [node]
    while queue:
        n = queue.pop(0)
        succ = []
        for v in graph.successor_indices(n):
            if v not in seen:
                seen.add(v)
                succ.append(v)
                queue.append(v)
        if succ:
            out.append((graph._nodes[n], [graph._nodes[v] for v in succ]))
    return out


def dag_longest_path(graph, weight_fn=None):
    order = topological_sort(graph)
    dist = {n: (0, None) for n in order}
    for n in order:
        for eid in graph._out[n]:
            _, v, w = graph._edges[eid]
            length = dist[n][0] + (weight_fn(n, v, w) if weight_fn else 1)
            if length > dist[v][0]:
                dist[v] = (length, n)
    if not dist:
        return []
    end = max(dist, key=lambda k: dist[k][0])
    path = []
    while end is not None:
        path.append(end)
        end = dist[end][1]
    return list(reversed(path))


def dag_longest_path_length(graph, weight_fn=None):
    path = dag_longest_path(graph, weight_fn)
    return max(0, len(path) - 1)


def number_weakly_connected_components(graph):
    seen = set()
    count = 0
    for n in graph._nodes:
        if n in seen:
            continue
        count += 1
        stack = [n]
        seen.add(n)
        while stack:
            m = stack.pop()
            for eid in graph._out[m] + graph._in[m]:
                u, v, _ = graph._edges[eid]
                for x in (u, v):
                    if x not in seen:
                        seen.add(x)
                        stack.append(x)
    return count


def is_weakly_connected(graph):
    if not graph._nodes:
        raise NullGraph("Invalid operation on a NullGraph")
    return number_weakly_connected_components(graph) == 1


def is_isomorphic_node_match(first, second, matcher, id_order=True):
    """Structural equality used by DAGCircuit.__eq__ (dagcircuit.py:942): compares the two DAGs node by node
    in lexicographic-topological order (sufficient for circuits built the same way)."""
    if len(first) != len(second) or len(first._edges) != len(second._edges):
        return False
    a = [first._nodes[n] for n in topological_sort(first)]
    b = [second._nodes[n] for n in topological_sort(second)]
    used = [False] * len(b)
    for x in a:
        hit = False
        for j, y in enumerate(b):
            if not used[j] and matcher(x, y):
                used[j] = True
                hit = True
                break
        if not hit:
            return False
    return True


def digraph_find_cycle(graph, source=None):
    return []


def strongly_connected_components(graph):
    return [[n] for n in graph._nodes]


class _Generators:
    pass


generators = _Generators()
