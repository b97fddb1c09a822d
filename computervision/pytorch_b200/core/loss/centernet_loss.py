"""Mirror of the one piece of reference core/loss/centernet_loss.py that the decode path uses:
`RegL1Loss.gather_feat` (:37-43).  The fused CenterNet kernel gathers reg / wh itself."""
import torch


class RegL1Loss:
    @staticmethod
    def gather_feat(feat, ind):
        """feat (B, H, W, C), ind (B, K) flat pixel indices -> (B, K, C)."""
        flat = feat.reshape(feat.size(0), -1, feat.size(3))
        idx = ind.unsqueeze(2).to(torch.int64).expand(-1, -1, flat.size(2))
        return torch.gather(flat, dim=1, index=idx)
