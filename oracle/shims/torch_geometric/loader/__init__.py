from oracle.thirdparty.pyg import Batch


class DataLoader:
    def __init__(self, dataset, batch_size=1, shuffle=False):
        assert not shuffle
        self.dataset, self.batch_size = list(dataset), batch_size

    def __iter__(self):
        for i in range(0, len(self.dataset), self.batch_size):
            yield Batch.from_data_list(self.dataset[i:i + self.batch_size])

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size
