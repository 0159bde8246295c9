"""Progressive rendering across the GPUs of one box: who renders which path indices, and how the per-rank
framebuffers come together.

Reference semantics being partitioned: view_render() covers path indices [counter, end) with end = counter + W*H per
progression (src/view.c:636-638); paths are independent given their index (render_sample_path(i), src/render.d/gi.c:81-88) and
the only shared state is the additive framebuffer.  Rank g of N renders progressions g, g+N, g+2N, ... -- the very index
ranges a 1-GPU run uses for those progressions, so the union over ranks after K steps is exactly progressions 0..K*N-1 --
and each progression's framebuffer is summed into rank 0 with one reduce (NCCL over NVLink on the box, gloo in the CPU tests).
Buffers are double-buffered: the reduce of progression s runs on a side stream while s+1 renders into the other buffer.
"""
import torch


def progression_range(step, rank, world, paths_per_progression):
    """(first_index, count) of the progression rank `rank` renders at its local step `step`"""
    return (step * world + rank) * paths_per_progression, paths_per_progression


class FramebufferReducer:
    """owns the per-rank accumulation buffers (H x W x 3 float32) and the root's running sum.

    On CUDA everything that follows a progression -- the reduce to rank 0, the root's accumulate, clearing the buffer for its
    next use and (optionally) mirroring the root's running sum into pinned HOST memory -- is queued on ONE side stream in that
    order; the render stream only waits for the event that marks its next buffer as cleared.  Nothing of it runs on the render
    stream or blocks the host, so a progression's epilogue overlaps the next progression's kernels."""

    def __init__(self, height, width, device, rank=0, world=1, dist=None, nbuf=2, host_mirror=None):
        self.rank, self.world, self.dist = rank, world, dist
        self.cuda = torch.device(device).type == "cuda"
        n = nbuf if world > 1 else 1
        self.bufs = [torch.zeros(height, width, 3, device=device) for _ in range(n)]
        self.accum = torch.zeros(height, width, 3, device=device) if (world > 1 and rank == 0) else None
        self.pending = [None] * n        # CPU path: the async work handle; CUDA path: the event after the buffer was cleared
        self.comm = torch.cuda.Stream() if (self.cuda and world > 1) else None
        self.host_mirror = host_mirror   # pinned (H, W, 3) float32 tensor on rank 0 or None
        self.submitted = 0

    def _retire(self, k):
        if self.pending[k] is None:
            return
        if self.comm is not None:        # CUDA: the side stream has done (or will do) the work; order the render stream behind it
            torch.cuda.current_stream().wait_event(self.pending[k])
        else:
            self.pending[k].wait()
            if self.accum is not None:
                self.accum.add_(self.bufs[k])
            self.bufs[k].zero_()
        self.pending[k] = None

    def acquire(self, step):
        """the buffer to render local step `step` into (ordered behind the reduce that last used it)"""
        k = step % len(self.bufs)
        self._retire(k)
        return self.bufs[k]

    def submit(self, step):
        """the buffer of local step `step` is complete on the current stream: start summing it into rank 0"""
        if self.world == 1:
            if self.host_mirror is not None:
                self.host_mirror.copy_(self.bufs[0], non_blocking=True)
            return
        k = step % len(self.bufs)
        if self.comm is not None:
            self.comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm):
                self.dist.reduce(self.bufs[k], 0, async_op=True).wait()     # stream-level wait: the host does not block
                if self.accum is not None:
                    self.accum.add_(self.bufs[k])
                    if self.host_mirror is not None:
                        self.host_mirror.copy_(self.accum, non_blocking=True)
                self.bufs[k].zero_()
                ev = torch.cuda.Event()
                ev.record(self.comm)
                self.pending[k] = ev
        else:
            self.pending[k] = self.dist.reduce(self.bufs[k], 0, async_op=True)
        self.submitted += 1

    def finish(self):
        """all reduces retired; returns the summed framebuffer on rank 0 (the local one when world == 1), None elsewhere"""
        for k in range(len(self.bufs)):
            self._retire(k)
        if self.world == 1:
            return self.bufs[0]
        return self.accum

    def clear(self):
        self.finish()
        for b in self.bufs:
            b.zero_()
        if self.accum is not None:
            self.accum.zero_()


def reduce_dbor(levels, rank=0, world=1, dist=None):
    """the `--dbor n` cascade buffers (src/view.c:497-522) are additive like the framebuffer, but only needed at export
    (view_write_images, view.c:553-556): one reduce of the stacked (n, H, W, 3) tensor to rank 0 at the end instead of one per
    progression.  Returns the sum on rank 0, None on the other ranks (the tensor itself when world == 1)."""
    if world == 1:
        return levels
    dist.reduce(levels, 0)
    return levels if rank == 0 else None

