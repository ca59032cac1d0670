-- TEST FIXTURE (tests/lua_interp.py runs it): the smallest stand-in for radio/core/block.lua that the glue in
-- lua/radio_b200/ touches -- classes made by factory(), instances with `inputs` / `outputs` port arrays whose ports carry
-- `owner`, `name`, `data_type` (and `pipes` on outputs), add_type_signature / differentiate / get_input_type /
-- get_output_type / get_rate.  Written for the test, not derived from the reference's implementation.
local M = {}

function M.Input(name, data_type) return {name = name, data_type = data_type} end
function M.Output(name, data_type) return {name = name, data_type = data_type} end

local Block = {}
Block.__index = Block

function Block:add_type_signature(inputs, outputs)
    self.inputs, self.outputs = {}, {}
    for i, p in ipairs(inputs) do self.inputs[i] = {owner = self, name = p.name, data_type = p.data_type} end
    for i, p in ipairs(outputs) do self.outputs[i] = {owner = self, name = p.name, data_type = p.data_type, pipes = {}} end
end
function Block:differentiate(input_types)
    self.differentiated = true
    for i, t in ipairs(input_types) do
        if self.inputs[i].data_type ~= t then error(self.name .. ": differentiate(): type mismatch on input " .. i) end
    end
end
function Block:get_input_type(i) return self.inputs[i or 1].data_type end
function Block:get_output_type(i) return self.outputs[i or 1].data_type end
function Block:get_rate() return self.rate end
function Block:instantiate() end
function Block:initialize() end
M.Block = Block

function M.factory(name, parent)
    local class = {}
    class.__index = class
    class.name = name
    setmetatable(class, {__index = parent or Block, __call = function (cls, ...)
        local self = setmetatable({}, cls)
        self:instantiate(...)
        return self
    end})
    return class
end

return M
