"""VAE skeleton.  Mirrors lib/networks/models.py (Local_Cond_RNVP_MC_Global_RNVP_VAE :13-258).

Encoder, latent prior flow and the base-Gaussian heads are latent-sized PyTorch; the per-point
decoder work is delegated to the CUDA flow stack (see flow_mixture.py for the K-component
`decode`).  Module / parameter / buffer names follow the reference so checkpoints interchange.
"""
import torch
import torch.nn as nn

from .decoders import GlobalRNVPDecoder, LocalCondRNVPDecoder
from .encoders import FeatureEncoder, PointNetCloudEncoder

_CFG_FIELDS = ('pc_enc_init_n_channels', 'pc_enc_init_n_features', 'pc_enc_n_features', 'g_latent_space_size',
               'g_prior_n_flows', 'g_prior_n_features', 'g_posterior_n_layers', 'p_latent_space_size',
               'p_prior_n_layers', 'p_decoder_n_flows', 'p_decoder_n_features', 'p_decoder_base_type',
               'p_decoder_base_var')


class Local_Cond_RNVP_MC_Global_RNVP_VAE(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.train_mode = kwargs.get('train_mode')
        self.mode = kwargs.get('util_mode')
        self.deterministic = kwargs.get('deterministic')
        for name in _CFG_FIELDS:
            setattr(self, name, kwargs.get(name))
        G, P3 = self.g_latent_space_size, self.p_latent_space_size

        self.pc_encoder = PointNetCloudEncoder(self.pc_enc_init_n_channels, self.pc_enc_init_n_features,
                                               self.pc_enc_n_features)
        self.g0_prior_mus = nn.Parameter(torch.empty(1, G))
        self.g0_prior_logvars = nn.Parameter(torch.empty(1, G))
        with torch.no_grad():
            nn.init.normal_(self.g0_prior_mus.data, mean=0.0, std=0.033)
            nn.init.normal_(self.g0_prior_logvars.data, mean=0.0, std=0.33)
        self.g_prior = GlobalRNVPDecoder(self.g_prior_n_flows, self.g_prior_n_features, G, weight_std=0.01)
        self.g_posterior = FeatureEncoder(self.g_posterior_n_layers, self.pc_enc_n_features[-1], G,
                                          deterministic=False, mu_weight_std=0.0033, mu_bias=0.0,
                                          logvar_weight_std=0.033, logvar_bias=0.0)
        base = self.p_decoder_base_type
        if base == 'free':
            self.p_prior = FeatureEncoder(self.p_prior_n_layers, G, P3, deterministic=False,
                                          mu_weight_std=0.001, mu_bias=0.0, logvar_weight_std=0.01, logvar_bias=0.0)
        elif base == 'freevar':
            self.register_buffer('p_prior_mus', torch.zeros((1, P3, 1)))
            self.p_prior = FeatureEncoder(self.p_prior_n_layers, G, P3, deterministic=True,
                                          mu_weight_std=0.01, mu_bias=0.0)
        elif base == 'fixed':
            self.register_buffer('p_prior_mus', torch.zeros((1, P3, 1)))
            self.register_buffer('p_prior_logvar', self.p_decoder_base_var * torch.ones((1, P3, 1)))
        self.pc_decoder = LocalCondRNVPDecoder(self.p_decoder_n_flows, self.p_decoder_n_features, G, weight_std=0.01)

    # ------------------------------------------------------------------ latent side
    def reparameterize(self, mu, logvar):
        std = torch.exp(0.5 * logvar)
        return torch.randn_like(std).mul(std).add_(mu)

    def encode(self, g_input):
        B, G = g_input.shape[0], self.g_latent_space_size
        out = {'g_prior_mus': [self.g0_prior_mus.expand(B, G)],
               'g_prior_logvars': [self.g0_prior_logvars.expand(B, G)]}
        if self.mode in ('training', 'autoencoding'):
            feats = torch.max(self.pc_encoder(g_input), dim=2)[0]
            out['g_posterior_mus'], out['g_posterior_logvars'] = self.g_posterior(feats)
            if self.mode == 'training':
                out['g_posterior_samples'] = self.reparameterize(out['g_posterior_mus'], out['g_posterior_logvars'])
            else:
                out['g_posterior_samples'] = out['g_posterior_mus']
            gs, mus, lvs = self.g_prior(out['g_posterior_samples'], mode='inverse')
            out['g_prior_samples'] = gs + [out['g_posterior_samples']]
        elif self.mode == 'generating':
            start = self.reparameterize(out['g_prior_mus'][0], out['g_prior_logvars'][0])
            gs, mus, lvs = self.g_prior(start, mode='direct')
            out['g_prior_samples'] = [start] + gs
        out['g_prior_mus'] += mus
        out['g_prior_logvars'] += lvs
        return out

    # ------------------------------------------------------------------ per-point side
    def base_gaussian(self, g_sample):
        """(mu_base, logvar_base), each (B,3)  -- models.py:169-193."""
        B, P3 = g_sample.shape[0], self.p_latent_space_size
        base = self.p_decoder_base_type
        if base == 'free':
            return self.p_prior(g_sample)
        if base == 'freevar':
            return self.p_prior_mus.view(1, P3).expand(B, P3), self.p_prior(g_sample)
        if base == 'fixed':
            return self.p_prior_mus.view(1, P3).expand(B, P3), self.p_prior_logvar.view(1, P3).expand(B, P3)
        raise ValueError('unknown p_decoder_base_type %r' % (base,))

    def one_flow_decode(self, p_input, g_sample, pc_decoder, n_sampled_points):
        """Single-decoder path with the reference's dict-of-lists contract (models.py:153-207)."""
        B, P3 = g_sample.shape[0], self.p_latent_space_size
        mu_b, lv_b = self.base_gaussian(g_sample)
        out = {'p_prior_mus': [mu_b.unsqueeze(2).expand(B, P3, n_sampled_points)],
               'p_prior_logvars': [lv_b.unsqueeze(2).expand(B, P3, n_sampled_points)]}
        if self.mode == 'training':
            ps, mus, lvs = pc_decoder(p_input, g_sample, mode='inverse')
            out['p_prior_samples'] = ps + [p_input]
        else:
            z = self.reparameterize(out['p_prior_mus'][0], out['p_prior_logvars'][0])
            ps, mus, lvs = pc_decoder(z, g_sample, mode='direct')
            out['p_prior_samples'] = [z] + ps
        out['p_prior_mus'] += mus
        out['p_prior_logvars'] += lvs
        return out

    def forward(self, g_input, p_input, images=None, n_sampled_points=None, labeled_samples=False, warmup=False):
        n_pts = p_input.shape[2] if n_sampled_points is None else n_sampled_points
        if images is not None and self.train_mode == 'p_rnvp_mc_g_rnvp_vae_ic':
            enc = self.encode(g_input, images)
        else:
            enc = self.encode(g_input)
        if self.mode in ('training', 'autoencoding'):
            g_sample = enc['g_posterior_samples']
        else:
            g_sample = enc['g_prior_samples'][-1]
        if labeled_samples:
            samples, labels, logits = self.decode(p_input, g_sample, n_pts, labeled_samples, warmup)
            return enc, samples, labels, logits
        out_dec, logits = self.decode(p_input, g_sample, n_pts, labeled_samples, warmup)
        return enc, out_dec, logits
