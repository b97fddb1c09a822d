"""Mirror of the one piece of reference core/loss/centernet_loss.py that the decode path uses:
`RegL1Loss.gather_feat` (:37-43).  The fused CenterNet kernel gathers reg / wh itself; this is the
standalone entry point, routed to cvpp_gather_feat."""
from ... import ops


class RegL1Loss:
    @staticmethod
    def gather_feat(feat, ind):
        """feat (B, H, W, C), ind (B, K) flat pixel indices (int32 / int64) -> (B, K, C)."""
        flat = feat.reshape(feat.size(0), -1, feat.size(3))
        return ops.gather_feat(flat.float(), ind)
