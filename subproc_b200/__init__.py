"""subproc_b200 -- B200-native batched Othello hot path behind the interfaces of ysnrkdm/subproc.

    subproc_b200.board        drop-in for the reference's ``board`` module (single game, CUDA-backed)
    subproc_b200.batched      BatchedOthello: Board semantics for B games in HBM
    subproc_b200.ops          one function per kernel (torch tensors -> C ABI)
    subproc_b200.game_runner  GameRunner / play_a_game over the lock-step playout kernel
    subproc_b200.parameter    ProgressPositionMovesParameter / counts() on the feature kernel
    subproc_b200.learner      per-phase regression from on-GPU statistics (+ NCCL all-reduce)
    subproc_b200.csrc         the CUDA sources; include/othello_b200.h is the C ABI

The CUDA library is the product.  Nothing here falls back to a CPU implementation.
"""
__version__ = "0.1.0"
