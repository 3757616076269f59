"""State wire format: byte-compatible with the reference's own `serialize_state`.

The reference pushes `env.serialize_state(state)` to every client on every step (match_server.py:206-207) and the
client rebuilds it with `server_environment.deserialize_state(...)` (ClientEnvironment.py:176-198) to call
`valid_actions` locally (:327-328).  Both are `dill.dumps` / `dill.loads` of the state object
(BlokusEnvironment.py:305-337, tictactoe_2p_env.py:185-217, TronGridEnvironment.py:434-463), i.e. a pickle stream in
which a Blokus state names the classes `colosseumrl.envs.blokus.board.Board` and `colosseumrl.envs.blokus.ai.AI`
(Tic Tac Toe and Tron states are tuples of numpy arrays / None: no class paths at all).

`dumps` writes exactly that stream for the adapters' look-alike objects: instances of the look-alike `Board` / `AI`
are emitted as instances of the REFERENCE's class paths (a GLOBAL opcode by name -- the reference package does not have
to be importable where the state is produced), with the attribute names the reference's methods read
(`board_contents`; `player_color`, `current_pieces`, `player_score`).  An untouched reference client therefore
unpickles real `Board` / `AI` objects and its `valid_actions` works on them.  `loads` reads the reference's own
streams (and ours): the two class paths resolve to the reference's classes when that package is importable and to the
look-alikes otherwise; anything else a stream names is resolved normally.
"""
import io
import pickle

REFERENCE_CLASS_PATHS = {"Board": ("colosseumrl.envs.blokus.board", "Board"), "AI": ("colosseumrl.envs.blokus.ai", "AI")}
_lookalikes = {}          # name -> look-alike class, registered by colosseumrl_b200.single


def register_lookalike(name, cls):
    _lookalikes[name] = cls
    return cls


class _ReferencePathPickler(pickle._Pickler):
    """Pure-Python pickler (protocol 2, the one dill.loads of any version reads) whose only change is how the two
    look-alike classes are named in the stream."""

    def save_global(self, obj, name=None):
        for key, cls in _lookalikes.items():
            if obj is cls:
                module, qual = REFERENCE_CLASS_PATHS[key]
                self.write(pickle.GLOBAL + module.encode("ascii") + b"\n" + qual.encode("ascii") + b"\n")
                self.memoize(obj)
                return
        super().save_global(obj, name)



def dumps(state) -> bytes:
    """`dill.dumps(state)`-compatible bytes of a reference-layout state."""
    buf = io.BytesIO()
    _ReferencePathPickler(buf, protocol=2).dump(state)
    return buf.getvalue()


class _ReferencePathUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        for key, (m, q) in REFERENCE_CLASS_PATHS.items():
            if module == m and name == q:
                try:                               # the real class when the reference package is there
                    return super().find_class(module, name)
                except (ImportError, AttributeError):
                    return _lookalikes[key]
        return super().find_class(module, name)


def loads(data):
    """`dill.loads`-compatible: reads the reference's own serialized states and the ones `dumps` writes."""
    return _ReferencePathUnpickler(io.BytesIO(bytes(data))).load()
