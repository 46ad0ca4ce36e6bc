"""CPU emulation of the tower operand precision on the shipped trained checkpoint: which of the three terms of
the hi/lo split (a_hi*w_hi + a_hi*w_lo + a_lo*w_hi) can be dropped, globally or for a single layer, within 1e-4?
fp64 accumulation, operands rounded to fp16 (hi) and fp16 residual (lo); BN folded like the device does."""
import sys, numpy as np, torch, torch.nn.functional as F
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pvnet_ref, omok_oracle as O
torch.set_num_threads(8)
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'trained_9x9_180927.npz'))
sd = {k: torch.from_numpy(z[k]) for k in z.files}
B=9
rs = np.random.RandomState(0)
ids = [(0,) + tuple(int(a) for a in rs.permutation(81)[:rs.randint(0, 50)]) for _ in range(256)]
x = torch.from_numpy(np.stack([O.get_state_pt(i, B, 5) for i in ids]).astype(np.float32))
pr, vr = pvnet_ref.pvnet_forward(sd, x)

def r16(t): return t.half().float()
def split(t):
    hi = r16(t); lo = r16(t - hi); return hi, lo

def fold(wname, bnname):
    w = sd[wname].float(); g=sd[bnname+'.weight']; b=sd[bnname+'.bias']; m=sd[bnname+'.running_mean']; v=sd[bnname+'.running_var']
    s = g/torch.sqrt(v+1e-5)
    return w*s[:,None,None,None], b-m*s

def conv(a, w, mode):
    # returns conv in float64 accumulate with operand roundings
    if mode=='fp32': return F.conv2d(a.double(), w.double(), padding=1).float()
    ah, al = split(a); wh, wl = split(w)
    d = lambda p,q: F.conv2d(p.double(), q.double(), padding=1)
    if mode=='both16': r = d(ah,wh)
    elif mode=='wsplit': r = d(ah,wh)+d(ah,wl)
    elif mode=='asplit': r = d(ah,wh)+d(al,wh)
    elif mode=='x3': r = d(ah,wh)+d(ah,wl)+d(al,wh)
    elif mode=='bf16x?': r=None
    return r.float()

def fwd(mode):
    w,b = fold('conv1.weight','bn1')
    h = F.relu(conv(x,w,mode)+b[None,:,None,None])
    for i in range(10):
        r = h
        w,b = fold(f'layers.{i}.conv1.weight', f'layers.{i}.bn1')
        o = F.relu(conv(h,w,mode)+b[None,:,None,None])
        w,b = fold(f'layers.{i}.conv2.weight', f'layers.{i}.bn2')
        o = conv(o,w,mode)+b[None,:,None,None]
        h = F.relu(o+r)
    p = F.relu(pvnet_ref._bn(F.conv2d(h, sd["policy_head.policy_head.weight"]), sd, "policy_head.policy_bn"))
    p = p.reshape(p.shape[0], -1)
    p = F.softmax(F.linear(p, sd["policy_head.policy_fc.weight"], sd["policy_head.policy_fc.bias"]), dim=-1)
    v = F.relu(pvnet_ref._bn(F.conv2d(h, sd["value_head.value_head.weight"]), sd, "value_head.value_bn"))
    v = v.reshape(v.shape[0], -1)
    v = F.relu(F.linear(v, sd["value_head.value_fc1.weight"], sd["value_head.value_fc1.bias"]))
    v = torch.tanh(F.linear(v, sd["value_head.value_fc2.weight"], sd["value_head.value_fc2.bias"]))
    return p, v.reshape(-1)
with torch.no_grad():
    for mode in ['fp32','both16','wsplit','asplit','x3']:
        p,v = fwd(mode)
        print(mode, 'dp %.3e dv %.3e'%((p-pr).abs().max().item(), (v-vr).abs().max().item()))

def fwd_modes(modes):
    w,b = fold('conv1.weight','bn1')
    h = F.relu(conv(x,w,modes[0])+b[None,:,None,None])
    for i in range(10):
        r = h
        w,b = fold(f'layers.{i}.conv1.weight', f'layers.{i}.bn1')
        o = F.relu(conv(h,w,modes[1+2*i])+b[None,:,None,None])
        w,b = fold(f'layers.{i}.conv2.weight', f'layers.{i}.bn2')
        o = conv(o,w,modes[2+2*i])+b[None,:,None,None]
        h = F.relu(o+r)
    p = F.relu(pvnet_ref._bn(F.conv2d(h, sd["policy_head.policy_head.weight"]), sd, "policy_head.policy_bn"))
    p = p.reshape(p.shape[0], -1)
    p = F.softmax(F.linear(p, sd["policy_head.policy_fc.weight"], sd["policy_head.policy_fc.bias"]), dim=-1)
    v = F.relu(pvnet_ref._bn(F.conv2d(h, sd["value_head.value_head.weight"]), sd, "value_head.value_bn"))
    v = v.reshape(v.shape[0], -1)
    v = F.relu(F.linear(v, sd["value_head.value_fc1.weight"], sd["value_head.value_fc1.bias"]))
    v = torch.tanh(F.linear(v, sd["value_head.value_fc2.weight"], sd["value_head.value_fc2.bias"]))
    return p, v.reshape(-1)
with torch.no_grad():
    for l in range(21):
        for alt in ['both16','wsplit','asplit']:
            modes = ['x3']*21; modes[l]=alt
            p,v = fwd_modes(modes)
            print(l, alt, 'dp %.3e dv %.3e'%((p-pr).abs().max().item(), (v-vr).abs().max().item()))
