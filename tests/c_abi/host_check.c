/* Compiled as plain C (gcc -std=c99 -pedantic) by tests/test_host_cpu.py: include/b2v.h must be a valid C header and
 * the host-only entry points must be callable from C.  No device work is done here. */
#include <stdio.h>
#include <string.h>

#include "b2v.h"

int main(void) {
  int64_t ts[64];
  int n, i;
  b2v_unet* u = NULL;
  b2v_unet_desc d;
  b2v_sampler_cfg cfg;
  if (b2v_abi_version() != B2V_ABI_VERSION) return 1;
  n = b2v_ddim_timesteps(1000, 50, ts, 64); /* inference/sampler.py:221-239: 999, 980, ..., 0 */
  if (n != 51 || ts[0] != 999 || ts[1] != 980 || ts[50] != 0) return 2;
  for (i = 1; i < n; ++i)
    if (ts[i] >= ts[i - 1]) return 3;
  if (b2v_ddim_timesteps(1000, 7, ts, 64) != 9 || ts[0] != 999 || ts[1] != 994) return 4; /* stride 142 + appended 999 */
  if (b2v_ddim_timesteps(1000, 50, ts, 8) >= 0 || strlen(b2v_last_error()) == 0) return 5;
  memset(&d, 0, sizeof d);
  d.latent_dim = 4, d.model_channels = 64, d.num_res_blocks = 1, d.num_levels = 2, d.channel_mult[0] = 1,
  d.channel_mult[1] = 2, d.attention_mask = 2, d.num_heads = 2, d.time_embed_dim = 128;
  memset(&cfg, 0, sizeof cfg);
  cfg.sampler = 0, cfg.n = n, cfg.timesteps = ts;
  /* without a B200 every create fails loudly (there is no CPU fallback); with one it succeeds */
  if (b2v_unet_create(&u, &d) != 0) {
    printf("create: %s\n", b2v_last_error());
    if (strlen(b2v_last_error()) == 0) return 6;
  } else {
    b2v_unet_destroy(u);
  }
  printf("ok abi=%d eps_mse_ws(4)=%lu launches=%lld\n", b2v_abi_version(), (unsigned long)b2v_eps_mse_ws_bytes(4),
         b2v_launch_count());
  return 0;
}
