import sys, os
sys.path.insert(0,'tests')
import driver_harness as H
root='/tmp/kgrec_synth2/'
H.make_dataset(root, users=300, items=1200, ratings=20000, entities=600, relations=8, triples=30000)
COMMON = ["-dataset", "ml1m", "-embedding_size", "32", "-batch_size", "256", "-seed", "3", "-num_processes", "2",
          "-optimizer_type", "Adagrad", "-learning_rate", "0.05", "-topn", "10", "-data_path", root]
for steps, iv in ((4, 2), (24, 12), (100,50)):
    fl=["-model_type","transr","-kg_test_files","valid.dat","-training_steps",str(steps),"-eval_interval_steps",str(iv)]
    a,_,_=H.run_driver("kg", COMMON+fl, '/tmp/kgrec_logd', 'ref', cpu=True)
    b,_,_=H.run_driver("kg", COMMON+fl, '/tmp/kgrec_logd', 'new', dropin=True)
    print(steps, iv, 'ref', a['train_loss'], a['kg']); print(steps, iv, 'new', b['train_loss'], b['kg'], flush=True)
