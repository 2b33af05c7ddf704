"""Execute the reference's OWN advantage-estimation loop (TEST INFRASTRUCTURE).

`Runner.run` of the reference (src/ppo_multi_agent_new.py:167-220) cannot be called here: the module imports
TensorFlow 1.x, which is not installed.  Its GAE tail (:205-218) however is 14 lines of pure numpy.  This module parses
the reference file with `ast`, takes those statements out of `Runner.run` UNMODIFIED (nothing is copied into the repo:
the nodes are compiled from the reference's file at run time) and executes them against a stand-in `self`.  That pins
oracle/gae_oracle.py -- and through it snk_gae -- on the reference's code instead of on a restatement.
"""
import ast
import os
import types

import numpy as np

import ref_loader


def _source_path():
    if ref_loader.source() == "tree":
        return os.path.join(ref_loader.REFERENCE_ROOT, "src", "ppo_multi_agent_new.py")
    return None  # the staged install (baseline/_ref) holds the env path only, not the learner


def available():
    p = _source_path()
    return p is not None and os.path.exists(p)


def _gae_statements():
    """The statements of Runner.run from `mb_returns = np.zeros_like(mb_rewards)` through `mb_returns = mb_advs + mb_values`."""
    path = _source_path()
    tree = ast.parse(open(path).read(), path)
    run = next(f for c in tree.body if isinstance(c, ast.ClassDef) and c.name == "Runner"
               for f in c.body if isinstance(f, ast.FunctionDef) and f.name == "run")
    def targets(node):
        return [t.id for t in getattr(node, "targets", []) if isinstance(t, ast.Name)]
    start = next(i for i, n in enumerate(run.body) if isinstance(n, ast.Assign) and targets(n) == ["mb_returns"])
    end = max(i for i, n in enumerate(run.body) if isinstance(n, ast.Assign) and targets(n) == ["mb_returns"])
    stmts = run.body[start:end + 1]
    assert any(isinstance(n, ast.For) for n in stmts), "the reversed GAE loop was not found"
    return stmts, (stmts[0].lineno, stmts[-1].end_lineno)


def reference_gae(rewards, values, dones, last_values, last_dones, gamma, lam):
    """(advs, returns) computed by the reference's statements.  Array dtypes as Runner.run prepares them (:198-203)."""
    stmts, _ = _gae_statements()
    code = compile(ast.Module(body=stmts, type_ignores=[]), _source_path(), "exec")
    mb_rewards = np.asarray(rewards, dtype=np.float32)
    ns = {
        "np": np, "self": types.SimpleNamespace(nsteps=mb_rewards.shape[0], dones=np.asarray(last_dones, dtype=bool),
                                                 gamma=gamma, lam=lam),
        "mb_rewards": mb_rewards, "mb_values": np.asarray(values, dtype=np.float32),
        "mb_dones": np.asarray(dones, dtype=bool), "last_values": np.asarray(last_values, dtype=np.float32),
    }
    exec(code, ns)
    return ns["mb_advs"], ns["mb_returns"]


def cited_lines():
    return _gae_statements()[1]
