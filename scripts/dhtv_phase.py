"""Phase timers of the DHTV kernels (PBB_LIB = a -DPBB_PHASE_TIMING build): one calculate_mapping at C3 size."""
import sys, torch
sys.path.insert(0, '.')
from oracle import synth
from pb_bss_b200.distribution import CACGMMTrainer
from pb_bss_b200.permutation_alignment import DHTVPermutationAlignment
F, T, K = 513, 500, 3
y, _ = synth.structured_stft(F, T, 8, K, seed=5)
init = synth.init_affiliation(F, K, T, seed=7)
yd = torch.from_numpy(y).cuda()
m = CACGMMTrainer().fit(yd, initialization=torch.from_numpy(init).cuda(), iterations=100)
mask = m.predict(yd).permute(1, 0, 2).contiguous()
al = DHTVPermutationAlignment.from_stft_size(1024)
al.calculate_mapping(mask)
print('---- second call ----', file=sys.stderr, flush=True)
al.calculate_mapping(mask)
