import os, torch, glob
print('allowed cpus:', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], '...')
for n in sorted(glob.glob('/sys/devices/system/node/node*')):
    try: print(n.split('/')[-1], open(n+'/cpulist').read().strip())
    except Exception as e: print(n, e)
for i in range(torch.cuda.device_count()):
    p=torch.cuda.get_device_properties(i)
    bdf=f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    try: node=open(f'/sys/bus/pci/devices/{bdf}/numa_node').read().strip()
    except Exception as e: node=str(e)
    print('gpu',i,bdf,'numa node',node)
