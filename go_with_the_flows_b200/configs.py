"""The reference's model configurations, restated as dicts (values of configs/*.yaml that the model
and loss constructors read; train_ae.py:47-65 splats the YAML into them and injects weights_type).
Synthetic-data runs (bench, tests) use these because the GPU box has no /root/reference."""

_COMMON = dict(
    deterministic=False, util_mode='training',
    pc_enc_init_n_channels=3, pc_enc_init_n_features=64, pc_enc_n_features=[128, 256, 512],
    g_posterior_n_layers=1, g_prior_n_flows=7,
    p_latent_space_size=3, p_prior_n_layers=1, p_decoder_n_flows=21, p_decoder_n_features=64,
    p_decoder_base_var=-3.9551, n_components=4, params_reduce_mode='depth_and_feature',
    weights_type='learned_weights', pnll_weight=1.0, gnll_weight=1.0, gent_weight=1.0,
    cloud_size=2048,
)

# config_generative_modeling_{airplane,car,chair}.yaml -> K=4, 11 triples, F=37, G=128, base 'free'
GENERATIVE = dict(_COMMON, train_mode='p_rnvp_mc_g_rnvp_vae', g_latent_space_size=128, g_prior_n_features=128,
                  p_decoder_base_type='free', batch_size=64)

# config_autoencoding.yaml -> K=4, 11 triples, F=33, G=512, base 'freevar'
AUTOENCODING = dict(_COMMON, train_mode='p_rnvp_mc_g_rnvp_vae', g_latent_space_size=512, g_prior_n_features=128,
                    p_decoder_base_type='freevar', p_decoder_base_var=-3.596, batch_size=128)

# config_SVR.yaml -> same decoder as autoencoding, image-conditioned prior, 2500 points
SVR = dict(_COMMON, train_mode='p_rnvp_mc_g_rnvp_vae_ic', g_latent_space_size=512, g_prior_n_features=128,
           g_prior_n_layers=1, p_decoder_base_type='freevar', p_decoder_base_var=0.0, batch_size=128,
           cloud_size=2500)

BY_NAME = {'generative': GENERATIVE, 'autoencoding': AUTOENCODING, 'svr': SVR}
