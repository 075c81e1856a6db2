"""CUDA-graph capture of whole render calls / training iterations.

Every buffer of the path is sized for the worst case and every data-dependent count lives in device memory, so a
render block (and a full forward + backward + Adam iteration) is a fixed launch sequence: ~170 (render) / ~400 (train)
launches are replayed as one graph and the host cost per step drops from milliseconds to tens of microseconds.
Opt-in: outputs of a graphed call are static buffers that the next call overwrites."""
import torch


def _tree_tensors(x):
    if torch.is_tensor(x):
        return [x]
    if isinstance(x, dict):
        return [t for v in x.values() for t in _tree_tensors(v)]
    if isinstance(x, (list, tuple)):
        return [t for v in x for t in _tree_tensors(v)]
    return []


class GraphedFn:
    """Captures `fn(**static_inputs)` once per input signature and replays it.

    inputs: dict name -> tensor (copied into static device buffers before each replay) or non-tensor (part of the key).
    fn must be free of host synchronisation and data-dependent host control flow."""

    def __init__(self, fn, device, warmup=3, state=None, capture_error_mode="global"):
        """state: tensors `fn` updates in place (parameters, optimizer moments, step counters).  They are snapshotted
        before the warm-up calls and restored after capture, so that building a graph has no side effect: the first
        replay is the first real application of `fn` (a captured optimizer step would otherwise be applied warmup + 1
        times to the first batch of every new input signature).
        capture_error_mode: "thread_local" when the captured region holds NCCL collectives (the process group's watchdog
        thread polls CUDA events, which a "global" capture would turn into an error)."""
        self.fn, self.device, self.warmup = fn, torch.device(device), warmup
        self.state = list(state) if state is not None else []
        self.capture_error_mode = capture_error_mode
        self.cache = {}

    @staticmethod
    def _key(inputs):
        k = []
        for name in sorted(inputs):
            v = inputs[name]
            k.append((name, tuple(v.shape), str(v.dtype)) if torch.is_tensor(v) else (name, repr(v)))
        return tuple(k)

    def __call__(self, **inputs):
        key = self._key(inputs)
        ent = self.cache.get(key)
        if ent is None:
            static = {n: (torch.empty(v.shape, dtype=v.dtype, device=self.device).copy_(v, non_blocking=True)
                          if torch.is_tensor(v) else v) for n, v in inputs.items()}
            snap = [t.detach().clone() for t in self.state]
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                for _ in range(self.warmup):
                    self.fn(**static)
            torch.cuda.current_stream(self.device).wait_stream(s)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode=self.capture_error_mode):
                out = self.fn(**static)
            with torch.no_grad():
                for t, v in zip(self.state, snap):
                    t.copy_(v)
            ent = self.cache[key] = (g, static, out)
        else:
            g, static, out = ent
            for n, v in inputs.items():
                if torch.is_tensor(v):
                    static[n].copy_(v, non_blocking=True)
        ent[0].replay()
        return ent[2]
